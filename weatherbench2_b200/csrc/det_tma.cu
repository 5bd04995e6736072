// K1 fast path -- TMA-staged persistent kernel (sm_100a).
//
// Same arithmetic and outputs as det_metrics_kernel (det_metrics.cu), but the
// slabs are moved HBM -> shared memory by the TMA engine (1-D bulk copies,
// `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes`, SASS
// UBLKCP) into a ring of stages guarded by full/empty mbarriers, so the number
// of bytes in flight per SM (~140 KB) no longer depends on registers or
// occupancy.  One persistent CTA per SM:
//   warp 8   : producer -- one elected lane walks the CTA's tile list (32-bit
//              incremental index math only) and issues 2 or 3 bulk copies per
//              tile (forecast, truth[, clim]);
//   warps 0-7: consumers -- warp pair q = w/2 owns tiles q, q+4, ...; each warp
//              of the pair reads one half of the tile with conflict-free
//              LDS.128, accumulates the unweighted f32 sums, releases the
//              stage, butterfly-reduces and lets lane r add region r's float64
//              weight * sums to its accumulators.
// A tile is one whole row (5760 B per operand at 1440 columns), so the producer
// issues one tile per ~750 cycles at full HBM rate.  Tiles are dealt to CTAs in
// contiguous ranges, so a CTA touches few fields; every (CTA, warp) writes
// float64 partials per touched field and a finalize kernel adds them in a
// fixed order (deterministic).
//
// Eligibility (checked by the caller): float32, every slab 16-byte aligned
// (offsets, row stride and ncol multiples of 4 elements), one column segment,
// no per-column / per-cell weight factor.  Everything else takes the LDG kernel.
#include "common.cuh"
#include "tma_utils.cuh"

namespace wb2 {

constexpr int kConsumerWarps = 8;
constexpr int kTmaThreads = (kConsumerWarps + 1) * 32;
constexpr int kTilesInFlight = kConsumerWarps / 2;  // tiles being consumed at once

struct TmaParams {
  const float* f;
  const float* t;
  const float* c;
  const int64_t* off_f;
  const int64_t* off_t;
  const int64_t* off_c;
  const double* row_w;  // [R][nrow]
  const double* seg_w;  // [R][1]
  double* partial;      // [ncta][kConsumerWarps][maxslots][R][WB2_DET_NSTAT]
  int64_t ntiles;       // nfield * tiles_per_field
  int32_t nrow, ncol;
  int64_t row_stride;
  int32_t nregion;
  int32_t zero_skip;
  int32_t tiles_per_field;  // = nrow (one tile per row)
  int32_t nstage;           // multiple of kTilesInFlight
  int32_t stage_op_bytes;   // bytes reserved per operand per stage (128-B mult.)
  int32_t maxslots;
};

template <bool CLIM, bool SKIPNA>
__device__ __forceinline__ void tma_cell(float f, float t, float c, float* acc) {
  constexpr int NSUM = CLIM ? 6 : 3;
  const float d = f - t;
  if (SKIPNA) {
    if (d == d) {
      acc[0] += d * d;
      acc[1] += fabsf(d);
      acc[2] += d;
      acc[NSUM] += 1.0f;
    }
  } else {
    acc[0] += d * d;
    acc[1] += fabsf(d);
    acc[2] += d;
  }
  if (CLIM) {
    const float fa = f - c, ta = t - c;  // weatherbench2/metrics.py:405-406
    const float v3 = fa * ta, v4 = fa * fa, v5 = ta * ta;
    if (SKIPNA) {
      if (v3 == v3) { acc[3] += v3; acc[NSUM + 1] += 1.0f; }
      if (fa == fa) { acc[4] += v4; acc[NSUM + 2] += 1.0f; }
      if (ta == ta) { acc[5] += v5; acc[NSUM + 3] += 1.0f; }
    } else {
      acc[3] += v3;
      acc[4] += v4;
      acc[5] += v5;
    }
  }
}

template <bool CLIM, bool SKIPNA>
__global__ void __launch_bounds__(kTmaThreads, 1) det_tma_kernel(const TmaParams p) {
  constexpr int NOPER = CLIM ? 3 : 2;
  constexpr int NSUM = CLIM ? 6 : 3;
  // without skipna the weight sum is the plain cell count of the tile
  constexpr int NCNT = SKIPNA ? (CLIM ? 4 : 1) : 0;
  constexpr int NS = NSUM + NCNT;

  extern __shared__ __align__(128) unsigned char smem[];
  const int nstage = p.nstage;
  const size_t stage_bytes = size_t(NOPER) * p.stage_op_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stage_bytes * nstage);
  uint64_t* empty = full + nstage;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstage; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2);  // both warps of the consuming pair
    }
    mbar_fence_init();
  }
  __syncthreads();

  // contiguous, balanced tile range of this CTA
  const int64_t per = p.ntiles / gridDim.x, extra = p.ntiles % gridDim.x;
  const int64_t b = blockIdx.x;
  const int64_t t0 = b * per + (b < extra ? b : extra);
  const int64_t t1 = t0 + per + (b < extra ? 1 : 0);
  const int64_t ncta_tiles = t1 - t0;
  const int64_t first_field = t0 / p.tiles_per_field;

  if (warp == kConsumerWarps) {
    // ------------------------------ producer --------------------------------
    if (lane == 0) {
      int64_t field = first_field;
      int row = static_cast<int>(t0 - first_field * p.tiles_per_field);
      const float* pf = p.f + p.off_f[field];
      const float* pt = p.t + p.off_t[field];
      const float* pc = CLIM ? p.c + p.off_c[field] : nullptr;
      const uint32_t bytes = static_cast<uint32_t>(p.ncol) * 4u;
      int s = 0;
      uint32_t use = 0;
      for (int64_t j = 0; j < ncta_tiles; ++j) {
        if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
        const int64_t e = int64_t(row) * p.row_stride;
        unsigned char* dst = smem + stage_bytes * s;
        mbar_arrive_expect_tx(&full[s], bytes * NOPER);
        tma_load_1d(dst, pf + e, bytes, &full[s]);
        tma_load_1d(dst + p.stage_op_bytes, pt + e, bytes, &full[s]);
        if (CLIM) tma_load_1d(dst + 2 * p.stage_op_bytes, pc + e, bytes, &full[s]);
        if (++s == nstage) { s = 0; ++use; }
        if (++row == p.nrow) {
          row = 0;
          ++field;
          if (j + 1 < ncta_tiles) {
            pf = p.f + p.off_f[field];
            pt = p.t + p.off_t[field];
            if (CLIM) pc = p.c + p.off_c[field];
          }
        }
      }
    }
    return;
  }

  // -------------------------------- consumers --------------------------------
  const int R = p.nregion;
  const bool zero_skip = p.zero_skip != 0;
  double accd[NS + 1];  // [NS] = weight sum when !SKIPNA
#pragma unroll
  for (int i = 0; i <= NS; ++i) accd[i] = 0.0;
  int64_t cur_field = -1;

  auto flush_field = [&](int64_t field) {
    // lane r holds region r's sums of `field`; write them in public layout
    if (field < 0) return;
    const int slot = static_cast<int>(field - first_field);
    if (lane < R) {
      double* out = p.partial +
                    (((int64_t(blockIdx.x) * kConsumerWarps + warp) * p.maxslots + slot) * R +
                     lane) * WB2_DET_NSTAT;
#pragma unroll
      for (int st = 0; st < WB2_DET_NSTAT; ++st) {
        double v = 0.0;
        if (st < 6) {
          if (st < NSUM) v = accd[st];
        } else {
          const int j = st - 6;
          if (SKIPNA) {
            if (j < NCNT) v = accd[NSUM + j];
          } else if (CLIM || j == 0) {
            v = accd[NS];
          }
        }
        out[st] = v;
      }
    }
#pragma unroll
    for (int i = 0; i <= NS; ++i) accd[i] = 0.0;
  };

  const int pair = warp >> 1, half = warp & 1;
  const int n4 = p.ncol >> 2;
  const int n4a = (n4 + 1) >> 1;
  const int i0 = half ? n4a : 0, i1 = half ? n4 : n4a;
  int64_t field = first_field;
  int row = static_cast<int>(t0 - first_field * p.tiles_per_field) + pair;
  int s = pair;
  uint32_t use = 0;
  for (int64_t j = pair; j < ncta_tiles; j += kTilesInFlight) {
    while (row >= p.nrow) { row -= p.nrow; ++field; }
    if (field != cur_field) {
      flush_field(cur_field);
      cur_field = field;
    }
    // region weight of this row (issued early, consumed after the tile)
    double w = 0.0;
    if (lane < R) w = p.row_w[int64_t(lane) * p.nrow + row] * p.seg_w[lane];

    const unsigned char* src = smem + stage_bytes * s;
    const float4* sf = reinterpret_cast<const float4*>(src);
    const float4* st = reinterpret_cast<const float4*>(src + p.stage_op_bytes);
    const float4* sc = reinterpret_cast<const float4*>(src + 2 * p.stage_op_bytes);

    float acc[NS > 0 ? NS : 1];
#pragma unroll
    for (int i = 0; i < NS; ++i) acc[i] = 0.0f;

    mbar_wait(&full[s], use & 1);
#pragma unroll 2
    for (int i = i0 + lane; i < i1; i += 32) {
      const float4 a = sf[i];
      const float4 bq = st[i];
      float4 cq = make_float4(0.f, 0.f, 0.f, 0.f);
      if (CLIM) cq = sc[i];
      tma_cell<CLIM, SKIPNA>(a.x, bq.x, cq.x, acc);
      tma_cell<CLIM, SKIPNA>(a.y, bq.y, cq.y, acc);
      tma_cell<CLIM, SKIPNA>(a.z, bq.z, cq.z, acc);
      tma_cell<CLIM, SKIPNA>(a.w, bq.w, cq.w, acc);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);  // this warp is done with the stage

#pragma unroll
    for (int i = 0; i < NS; ++i) acc[i] = warp_sum(acc[i]);
    if (lane < R && !(zero_skip && w == 0.0)) {
#pragma unroll
      for (int i = 0; i < NS; ++i) accd[i] += w * double(acc[i]);
      if (!SKIPNA) accd[NS] += w * double((i1 - i0) * 4);
    }
    row += kTilesInFlight;
    s += kTilesInFlight;
    if (s >= nstage) { s -= nstage; ++use; }
  }
  flush_field(cur_field);
}

__global__ void det_tma_finalize_kernel(const double* __restrict__ partial,
                                        double* __restrict__ out, int64_t ntiles, int ncta,
                                        int tiles_per_field, int maxslots, int per_field,
                                        int nwarps) {
  const int64_t field = blockIdx.x;
  const int64_t tf0 = field * tiles_per_field, tf1 = tf0 + tiles_per_field;
  const int64_t per = ntiles / ncta, extra = ntiles % ncta;
  for (int i = threadIdx.x; i < per_field; i += blockDim.x) {
    double v = 0.0;
    for (int b = 0; b < ncta; ++b) {
      const int64_t t0 = int64_t(b) * per + (b < extra ? b : extra);
      const int64_t t1 = t0 + per + (b < extra ? 1 : 0);
      if (t1 <= tf0 || t0 >= tf1) continue;
      const int slot = static_cast<int>(field - t0 / tiles_per_field);
      for (int w = 0; w < nwarps; ++w)
        v += partial[((int64_t(b) * nwarps + w) * maxslots + slot) * per_field + i];
    }
    out[field * per_field + i] = v;
  }
}

// Host wrapper so that other translation units (det_tma_seg.cu) can run the
// fixed-order finalize without relocatable device code.
int launch_det_tma_finalize(wb2_ctx* ctx, const double* partial, double* out, int64_t nfield,
                            int64_t ntiles, int ncta, int tiles_per_field, int maxslots,
                            int per_field, int nwarps) {
  det_tma_finalize_kernel<<<static_cast<unsigned>(nfield), 128, 0, ctx->stream>>>(
      partial, out, ntiles, ncta, tiles_per_field, maxslots, per_field, nwarps);
  WB2_CUDA_TRY(cudaGetLastError());
  return WB2_OK;
}

// Returns 1 if the TMA path ran, 0 if the launch is not eligible, < 0 on error.
int det_metrics_tma(wb2_ctx* ctx, bool clim, const void* f, const void* t, const void* c,
                    int64_t nfield, const int64_t* d_off_f, const int64_t* d_off_t,
                    const int64_t* d_off_c, const double* d_row_w, const double* d_seg_w,
                    const wb2_weights* w, int skipna, double* out) {
  const int noper = clim ? 3 : 2;
  if (w->ncol * 4 < 512) return 0;
  const int stage_op_bytes = (w->ncol * 4 + 127) / 128 * 128;
  const size_t budget = 220 * 1024;  // of the 227 KB a CTA may use on sm_100
  int nstage = static_cast<int>(budget / (size_t(noper) * stage_op_bytes));
  nstage = nstage / kTilesInFlight * kTilesInFlight;
  if (nstage > 32) nstage = 32;
  if (nstage < 2 * kTilesInFlight) return 0;
  const size_t smem = size_t(nstage) * noper * stage_op_bytes + 2 * nstage * sizeof(uint64_t);

  TmaParams p;
  p.f = static_cast<const float*>(f);
  p.t = static_cast<const float*>(t);
  p.c = static_cast<const float*>(c);
  p.off_f = d_off_f; p.off_t = d_off_t; p.off_c = d_off_c;
  p.row_w = d_row_w; p.seg_w = d_seg_w;
  p.nrow = w->nrow; p.ncol = w->ncol; p.row_stride = w->row_stride;
  p.nregion = w->nregion; p.zero_skip = w->zero_skip;
  p.tiles_per_field = w->nrow;
  p.ntiles = nfield * p.tiles_per_field;
  p.nstage = nstage; p.stage_op_bytes = stage_op_bytes;
  int ncta = ctx->num_sms;
  if (p.ntiles < ncta) ncta = static_cast<int>(p.ntiles);
  const int64_t per = (p.ntiles + ncta - 1) / ncta;
  p.maxslots = static_cast<int>((per + p.tiles_per_field - 1) / p.tiles_per_field) + 1;
  const size_t per_field = size_t(w->nregion) * WB2_DET_NSTAT;
  const size_t part_bytes = size_t(ncta) * kConsumerWarps * p.maxslots * per_field * sizeof(double);
  if (part_bytes > ctx->tma_partial_cap) {
    if (ctx->tma_partial) {
      WB2_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      WB2_CUDA_TRY(cudaFree(ctx->tma_partial));
      ctx->tma_partial = nullptr;
      ctx->tma_partial_cap = 0;
    }
    WB2_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&ctx->tma_partial), part_bytes));
    ctx->tma_partial_cap = part_bytes;
  }
  p.partial = ctx->tma_partial;
  WB2_CUDA_TRY(cudaMemsetAsync(p.partial, 0, part_bytes, ctx->stream));

  auto go = [&](auto kernel) -> int {
    WB2_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    kernel<<<ncta, kTmaThreads, smem, ctx->stream>>>(p);
    WB2_CUDA_TRY(cudaGetLastError());
    return WB2_OK;
  };
  int rc;
  if (clim) rc = skipna ? go(det_tma_kernel<true, true>) : go(det_tma_kernel<true, false>);
  else rc = skipna ? go(det_tma_kernel<false, true>) : go(det_tma_kernel<false, false>);
  if (rc != WB2_OK) return rc;
  det_tma_finalize_kernel<<<static_cast<unsigned>(nfield), 128, 0, ctx->stream>>>(
      p.partial, out, p.ntiles, ncta, p.tiles_per_field, p.maxslots,
      static_cast<int>(per_field), kConsumerWarps);
  WB2_CUDA_TRY(cudaGetLastError());
  ctx->launches += 2;
  return 1;
}

}  // namespace wb2
