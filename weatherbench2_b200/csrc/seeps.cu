// K9 -- SEEPS (Stable Equitable Error in Probability Space) maps (sm_100a).
//
// Replaces SpatialSEEPS.compute_chunk (weatherbench2/metrics.py:417-513): both
// forecast and truth precipitation are put into the categories dry / light /
// heavy with the reference's own comparisons (:449-454)
//     dry = x < dry_thr;  light = x > dry_thr and x < wet;  heavy = x >= wet
// (a value equal to dry_thr is in no category, and with wet < dry_thr a value
// can be in two: the score is the double sum  sum_fc sum_tc F[fc] T[tc] S[fc][tc]
// exactly like the reference's contingency-table dot product, :476-500), the
// 3 x 3 scoring matrix comes from the climatological dry fraction p1 of the
// cell (:482-492), and cells with p1 outside (min_p1, max_p1) are NaN
// (:503-504).  Like K6 the kernel can average `ngroup` time steps into one
// output map (Metric.compute, :117-138).  SEEPS itself (the spatial mean with
// skipna = True, :516-528) is K1 applied to these maps.
//
// One thread per grid cell of an output map; 8 B read (+ the wet-threshold
// slab, usually L2-resident) and 4 / ngroup bytes written per cell.
#include "common.cuh"

namespace wb2 {

constexpr int kSeepsThreads = 256;

struct SeepsParams {
  const float* f;
  const float* t;
  const float* wet;   // climatological wet threshold slabs
  const float* p1;    // [nrow][ncol] mean dry fraction
  float* out;         // [nout][nrow][ncol]
  const int64_t* off_f;   // [nout][ngroup]
  const int64_t* off_t;
  const int64_t* off_wf;  // wet threshold at the forecast's valid time
  const int64_t* off_wt;  // ... at the truth's valid time
  int64_t row_stride, wet_row_stride, cells_per_map;
  int32_t ngroup, nrow, ncol, bpm;
  float dry, min_p1, max_p1;
};

// category membership flags (bit 0 dry, 1 light, 2 heavy); NaN -> -1
__device__ __forceinline__ int seeps_cat(float x, float dry, float wet) {
  if (!(x == x)) return -1;
  int c = 0;
  if (x < dry) c |= 1;
  if (x > dry && x < wet) c |= 2;
  if (x >= wet) c |= 4;
  return c;
}

template <bool SKIPNA>
__global__ void __launch_bounds__(kSeepsThreads) seeps_maps_kernel(const SeepsParams p) {
  const int64_t j = blockIdx.x / p.bpm;
  const int64_t ci = int64_t(blockIdx.x % p.bpm) * kSeepsThreads + threadIdx.x;
  if (ci >= p.cells_per_map) return;
  const int row = static_cast<int>(ci / p.ncol);
  const int col = static_cast<int>(ci % p.ncol);
  const int64_t cell = int64_t(row) * p.row_stride + col;
  const int64_t wcell = int64_t(row) * p.wet_row_stride + col;
  const float nanv = __int_as_float(0x7fc00000);
  const double p1 = double(p.p1[ci]);
  const bool masked = !(p1 < double(p.max_p1)) || !(p1 > double(p.min_p1));  // :503-504
  // 0.5 * scoring matrix, S[forecast_cat][truth_cat] (:482-494)
  double S[3][3];
  S[0][0] = 0.0;                                   S[0][1] = 0.5 / (1.0 - p1);  S[0][2] = 2.0 / (1.0 - p1);
  S[1][0] = 0.5 / p1;                              S[1][1] = 0.0;               S[1][2] = 1.5 / (1.0 - p1);
  S[2][0] = 0.5 * (1.0 / p1 + 3.0 / (2.0 + p1));   S[2][1] = 1.5 / (2.0 + p1);  S[2][2] = 0.0;
  double acc = 0.0;
  int cnt = 0;
  float last = 0.f;
  for (int g = 0; g < p.ngroup; ++g) {
    const int64_t field = j * p.ngroup + g;
    const float f = ldg_stream(p.f + p.off_f[field] + cell);
    const float t = ldg_stream(p.t + p.off_t[field] + cell);
    const float wf = __ldg(p.wet + p.off_wf[field] + wcell);
    const float wt = __ldg(p.wet + p.off_wt[field] + wcell);
    const int cf = seeps_cat(f, p.dry, wf);
    const int ct = seeps_cat(t, p.dry, wt);
    float v;
    if (cf < 0 || ct < 0 || masked) {
      v = nanv;
    } else {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
          if ((cf >> a & 1) && (ct >> b & 1)) s += S[a][b];
      v = float(s);
    }
    last = v;
    if (SKIPNA) {
      if (v == v) { acc += double(v); ++cnt; }
    } else {
      acc += double(v);
    }
  }
  float r;
  if (p.ngroup == 1) r = last;
  else if (SKIPNA) r = cnt > 0 ? float(acc / double(cnt)) : nanv;
  else r = float(acc / double(p.ngroup));
  p.out[j * p.cells_per_map + ci] = r;
}

}  // namespace wb2

using namespace wb2;

extern "C" int wb2_seeps_maps(wb2_ctx* ctx, const float* f, const float* t, const float* wet,
                              const float* p1, int64_t nout, int32_t ngroup,
                              const int64_t* off_f, const int64_t* off_t,
                              const int64_t* off_wet_f, const int64_t* off_wet_t, int32_t nrow,
                              int32_t ncol, int64_t row_stride, int64_t wet_row_stride,
                              float dry_threshold, float min_p1, float max_p1, int skipna,
                              float* out) {
  WB2_NVTX("wb2_seeps_maps");
  WB2_REQUIRE(ctx != nullptr, "ctx is NULL");
  WB2_REQUIRE(nrow > 0 && ncol > 0 && row_stride >= ncol && wet_row_stride >= ncol,
              "bad grid: nrow=%d ncol=%d", nrow, ncol);
  WB2_REQUIRE(nout >= 0 && ngroup >= 1, "nout must be >= 0 and ngroup >= 1");
  if (nout == 0) return WB2_OK;
  WB2_REQUIRE(f && t && wet && p1 && out && off_f && off_t && off_wet_f && off_wet_t,
              "NULL argument");
  DeviceGuard guard(ctx->device);
  const int64_t nfield = nout * ngroup;
  Packer pk(ctx);
  const size_t o0 = pk.add(off_f, nfield * sizeof(int64_t));
  const size_t o1 = pk.add(off_t, nfield * sizeof(int64_t));
  const size_t o2 = pk.add(off_wet_f, nfield * sizeof(int64_t));
  const size_t o3 = pk.add(off_wet_t, nfield * sizeof(int64_t));
  WB2_TRY(pk.commit());
  SeepsParams p;
  p.f = f; p.t = t; p.wet = wet; p.p1 = p1; p.out = out;
  p.off_f = pk.dev<int64_t>(o0); p.off_t = pk.dev<int64_t>(o1);
  p.off_wf = pk.dev<int64_t>(o2); p.off_wt = pk.dev<int64_t>(o3);
  p.row_stride = row_stride; p.wet_row_stride = wet_row_stride;
  p.cells_per_map = int64_t(nrow) * ncol;
  p.ngroup = ngroup; p.nrow = nrow; p.ncol = ncol;
  p.bpm = static_cast<int32_t>((p.cells_per_map + kSeepsThreads - 1) / kSeepsThreads);
  p.dry = dry_threshold; p.min_p1 = min_p1; p.max_p1 = max_p1;
  WB2_REQUIRE(nout * int64_t(p.bpm) < (int64_t(1) << 31), "launch too large");
  const dim3 grid(static_cast<unsigned>(nout * p.bpm));
  if (skipna) seeps_maps_kernel<true><<<grid, kSeepsThreads, 0, ctx->stream>>>(p);
  else seeps_maps_kernel<false><<<grid, kSeepsThreads, 0, ctx->stream>>>(p);
  WB2_CUDA_TRY(cudaGetLastError());
  ctx->launches += 1;
  WB2_TRY(pk.release());
  return WB2_OK;
}
