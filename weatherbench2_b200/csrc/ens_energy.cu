// K3 -- energy-score sums (sm_100a).
//
// EnergyScoreSkill  = mean_m sqrt(SA((x_m - t)^2))           (metrics.py:1503-1517)
// EnergyScoreSpread = mean_{m<M-1} sqrt(SA((x_m - x_{m+1})^2)) (metrics.py:1471-1496)
// need 2M - 1 weighted spatial sums per field.  The reference materialises an
// M-sized difference array for each; running K1 once per member pair reads the
// ensemble ~4 times.  Here every member is read ONCE: a lane walks the members
// of its grid point keeping the previous member in a register and adds the two
// squared differences to per-lane accumulators (2 * MP - 1 of them), then the
// usual per-(row, segment) butterfly; lane (s mod 32) owns statistic s and
// applies the float64 region weights.
//
// Roofline: (4 M + 4) bytes per grid point, ~4 FP32 instructions per member:
// HBM-bound.  NaN propagates (skipna = False only; the skipna path keeps using
// K1 on member views).
#include "common.cuh"

namespace wb2 {

constexpr int kEsWarps = 4;
constexpr int kEsThreads = kEsWarps * 32;
constexpr int kEsMaxRegions = 4;

struct EsParams {
  const float* x;
  const float* t;
  const int64_t* off_x;
  const int64_t* off_t;
  const double* row_w;
  const int32_t* seg_start;
  const double* seg_w;
  const float* col_w;
  const float* cell_w;
  double* partial;  // [nfield][nblk][R][2 * MP]   (skill[MP], spread[MP-1], wsum)
  int64_t member_stride, row_stride;
  int32_t nmember, nrow, ncol, nregion, nseg, zero_skip, rows_per_block, nblk;
};

template <int MP>
__global__ void __launch_bounds__(kEsThreads, 3) energy_kernel(const EsParams p) {
  constexpr int NACC = 2 * MP;          // skill[0..MP), spread[0..MP-1), weight sum
  constexpr int NOWN = (NACC + 31) / 32;  // statistics owned per lane
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* red = reinterpret_cast<double*>(smem_raw);  // [warps][R][NACC]
  float* s_colw = reinterpret_cast<float*>(red + kEsWarps * kEsMaxRegions * NACC);

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t field = blockIdx.x / p.nblk;
  const int blk = blockIdx.x % p.nblk;
  const int R = p.nregion;
  const int M = p.nmember;
  const bool weighted = p.col_w != nullptr || p.cell_w != nullptr;
  if (p.col_w) {
    for (int i = threadIdx.x; i < p.ncol; i += kEsThreads) s_colw[i] = p.col_w[i];
    __syncthreads();
  }
  const float* __restrict__ px = p.x + p.off_x[field];
  const float* __restrict__ pt = p.t + p.off_t[field];
  const bool zero_skip = p.zero_skip != 0;

  double accd[NOWN][kEsMaxRegions];
#pragma unroll
  for (int i = 0; i < NOWN; ++i)
#pragma unroll
    for (int r = 0; r < kEsMaxRegions; ++r) accd[i][r] = 0.0;

  const int row0 = blk * p.rows_per_block;
  const int row1 = min(p.nrow, row0 + p.rows_per_block);
  for (int row = row0 + warp; row < row1; row += kEsWarps) {
    const int64_t rbase = int64_t(row) * p.row_stride;
    for (int k = 0; k < p.nseg; ++k) {
      const int s = p.seg_start[k];
      const int e = p.seg_start[k + 1];
      float acc[NACC];
#pragma unroll
      for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
      for (int col = s + lane; col < e; col += 32) {
        float wc = 1.f;
        if (weighted) {
          if (p.col_w) wc *= s_colw[col];
          if (p.cell_w) wc *= p.cell_w[int64_t(row) * p.ncol + col];
          if (zero_skip && wc == 0.f) continue;
        }
        const float* src = px + rbase + col;
        const float t = ldg_stream(pt + rbase + col);
        // members in chunks of CH loads in flight; only the previous member is
        // carried across chunks (keeps the register footprint ~2 MP + CH)
        constexpr int CH = MP <= 16 ? MP : (MP % 10 == 0 ? 10 : 8);
        float prev = 0.f;
#pragma unroll
        for (int m0 = 0; m0 < MP; m0 += CH) {
          float v[CH];
#pragma unroll
          for (int j = 0; j < CH; ++j)
            v[j] = (m0 + j) < M ? ldg_stream(src + int64_t(m0 + j) * p.member_stride) : 0.f;
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int m = m0 + j;
            if (m < M) {
              const float d = v[j] - t;
              acc[m] = fmaf(weighted ? wc * d : d, d, acc[m]);
              if (m > 0) {
                const float g = prev - v[j];
                acc[MP + m - 1] = fmaf(weighted ? wc * g : g, g, acc[MP + m - 1]);
              }
              prev = v[j];
            }
          }
        }
        acc[NACC - 1] += wc;
      }
#pragma unroll
      for (int i = 0; i < NACC; ++i) acc[i] = warp_sum(acc[i]);
      // lane (i mod 32) owns statistic i
#pragma unroll
      for (int r = 0; r < kEsMaxRegions; ++r) {
        if (r < R) {
          const double w = p.row_w[int64_t(r) * p.nrow + row] * p.seg_w[r * p.nseg + k];
          if (!(zero_skip && w == 0.0)) {
#pragma unroll
            for (int i = 0; i < NACC; ++i)
              if ((i & 31) == lane) accd[i >> 5][r] += w * double(acc[i]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NACC; ++i)
    if ((i & 31) == lane)
#pragma unroll
      for (int r = 0; r < kEsMaxRegions; ++r)
        if (r < R) red[(warp * kEsMaxRegions + r) * NACC + i] = accd[i >> 5][r];
  __syncthreads();
  double* out = p.partial + (field * p.nblk + blk) * int64_t(R) * NACC;
  for (int idx = threadIdx.x; idx < R * NACC; idx += kEsThreads) {
    const int r = idx / NACC, i = idx - r * NACC;
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kEsWarps; ++w) v += red[(w * kEsMaxRegions + r) * NACC + i];
    out[idx] = v;
  }
}

// partial [nfield][nblk][R][2*MP] -> out [nfield][R][4][M]
__global__ void energy_finalize_kernel(const double* __restrict__ partial,
                                       double* __restrict__ out, int nblk, int R, int MP, int M) {
  const int64_t field = blockIdx.x;
  const int nacc = 2 * MP;
  for (int idx = threadIdx.x; idx < R * 4 * M; idx += blockDim.x) {
    const int r = idx / (4 * M);
    const int q = (idx / M) % 4;
    const int m = idx % M;
    int src = -1;
    if (q == 0) src = m;
    else if (q == 1) src = m < M - 1 ? MP + m : -1;
    else src = (q == 3 && m >= M - 1) ? -1 : nacc - 1;  // weight sums
    double v = 0.0;
    if (src >= 0)
      for (int b = 0; b < nblk; ++b)
        v += partial[((field * nblk + b) * R + r) * int64_t(nacc) + src];
    out[field * int64_t(R) * 4 * M + idx] = v;
  }
}

template <int MP>
static int launch_energy(wb2_ctx* ctx, const EsParams& p, int64_t nfield) {
  const size_t smem = size_t(kEsWarps) * kEsMaxRegions * 2 * MP * sizeof(double) +
                      (p.col_w ? size_t(p.ncol) * sizeof(float) : 0);
  if (smem > 48 * 1024)
    WB2_CUDA_TRY(cudaFuncSetAttribute(energy_kernel<MP>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
  energy_kernel<MP><<<static_cast<unsigned>(nfield * p.nblk), kEsThreads, smem, ctx->stream>>>(p);
  WB2_CUDA_TRY(cudaGetLastError());
  return WB2_OK;
}

}  // namespace wb2

using namespace wb2;

extern "C" int wb2_energy_score(wb2_ctx* ctx, const void* x, const void* t, int dtype,
                                int32_t nmember, int64_t member_stride, int64_t nfield,
                                const int64_t* off_x, const int64_t* off_t,
                                const wb2_weights* w, double* out) {
  WB2_REQUIRE(ctx != nullptr, "ctx is NULL");
  WB2_REQUIRE(dtype == WB2_F32, "wb2_energy_score: only WB2_F32 inputs are supported");
  WB2_REQUIRE(nmember >= 1 && nmember <= 64,
              "wb2_energy_score: 1..64 ensemble members are supported (got %d)", nmember);
  WB2_TRY(validate_weights(w));
  WB2_REQUIRE(w->nregion <= kEsMaxRegions,
              "wb2_energy_score: at most %d regions per launch (got %d)", kEsMaxRegions,
              w->nregion);
  WB2_REQUIRE(out != nullptr, "out is NULL");
  WB2_REQUIRE(nfield >= 0 && nfield <= (int64_t(1) << 24), "nfield out of range");
  if (nfield == 0) return WB2_OK;
  WB2_REQUIRE(x && t && off_x && off_t, "x/t and their offset tables must not be NULL");
  DeviceGuard guard(ctx->device);

  int mp;
  if (nmember <= 2) mp = 2;
  else if (nmember <= 4) mp = 4;
  else if (nmember <= 8) mp = 8;
  else if (nmember <= 16) mp = 16;
  else if (nmember <= 32) mp = 32;
  else if (nmember <= 50) mp = 50;
  else mp = 64;
  int rows_per_block = 2 * kEsWarps;
  int nblk = (w->nrow + rows_per_block - 1) / rows_per_block;
  while (nblk * nfield < 4 * ctx->num_sms && rows_per_block > kEsWarps) {
    rows_per_block /= 2;
    nblk = (w->nrow + rows_per_block - 1) / rows_per_block;
  }
  const int R = w->nregion;
  Packer pk(ctx);
  size_t o_x = pk.add(off_x, nfield * sizeof(int64_t));
  size_t o_t = pk.add(off_t, nfield * sizeof(int64_t));
  size_t o_rw = pk.add(w->row_w, size_t(R) * w->nrow * sizeof(double));
  size_t o_ss = pk.add(w->seg_start, size_t(w->nseg + 1) * sizeof(int32_t));
  size_t o_sw = pk.add(w->seg_w, size_t(R) * w->nseg * sizeof(double));
  size_t o_cw = w->col_w ? pk.add(w->col_w, size_t(w->ncol) * sizeof(float)) : 0;
  size_t o_part = pk.reserve(size_t(nfield) * nblk * R * 2 * mp * sizeof(double));
  WB2_TRY(pk.commit());
  EsParams p;
  p.x = static_cast<const float*>(x);
  p.t = static_cast<const float*>(t);
  p.off_x = pk.dev<int64_t>(o_x); p.off_t = pk.dev<int64_t>(o_t);
  p.row_w = pk.dev<double>(o_rw); p.seg_start = pk.dev<int32_t>(o_ss);
  p.seg_w = pk.dev<double>(o_sw);
  p.col_w = w->col_w ? pk.dev<float>(o_cw) : nullptr;
  p.cell_w = w->cell_w;
  p.partial = pk.dev<double>(o_part);
  p.member_stride = member_stride; p.row_stride = w->row_stride;
  p.nmember = nmember; p.nrow = w->nrow; p.ncol = w->ncol;
  p.nregion = R; p.nseg = w->nseg; p.zero_skip = w->zero_skip;
  p.rows_per_block = rows_per_block; p.nblk = nblk;
  int rc;
  switch (mp) {
    case 2: rc = launch_energy<2>(ctx, p, nfield); break;
    case 4: rc = launch_energy<4>(ctx, p, nfield); break;
    case 8: rc = launch_energy<8>(ctx, p, nfield); break;
    case 16: rc = launch_energy<16>(ctx, p, nfield); break;
    case 32: rc = launch_energy<32>(ctx, p, nfield); break;
    case 50: rc = launch_energy<50>(ctx, p, nfield); break;
    default: rc = launch_energy<64>(ctx, p, nfield); break;
  }
  if (rc != WB2_OK) return rc;
  energy_finalize_kernel<<<static_cast<unsigned>(nfield), 128, 0, ctx->stream>>>(
      p.partial, out, nblk, R, mp, nmember);
  WB2_CUDA_TRY(cudaGetLastError());
  ctx->launches += 2;
  WB2_TRY(pk.release());
  return WB2_OK;
}
