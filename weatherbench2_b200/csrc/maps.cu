// K6 -- map-output ("Spatial*") deterministic metrics with the time mean fused
// in (sm_100a).
//
// Replaces SpatialBias / SpatialMSE / SpatialMAE .compute_chunk
// (weatherbench2/metrics.py:304-374): f - t, (f - t)^2, |f - t| per grid cell,
// and -- when ngroup > 1 -- the `.mean(time, skipna)` of Metric.compute
// (metrics.py:117-138) in the same pass, so the per-time maps are never
// written: (8 + 4 / ngroup) bytes per input cell instead of 8 + 4 + 4 + 4 / ngroup.
//
// One thread owns VEC consecutive cells of one output map and walks the ngroup
// (forecast, truth) slabs that average into it with 128-bit streaming loads;
// accumulation is float64 in registers in a fixed order (deterministic).
#include "common.cuh"

namespace wb2 {

constexpr int kMapThreads = 256;

struct MapParams {
  const void* f;
  const void* t;
  void* out;
  const int64_t* off_f;  // [nout][ngroup]
  const int64_t* off_t;
  int64_t row_stride;
  int64_t cells_per_map;  // nrow * ncol
  int32_t ngroup, nrow, ncol, ncv;  // ncv = ncol / VEC
  int32_t bpm;                      // blocks per map
};

template <typename T, int VEC> struct VecLoad;
template <> struct VecLoad<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    const float4 r = ldg_stream(reinterpret_cast<const float4*>(p));
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct VecLoad<float, 1> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[1]) { v[0] = ldg_stream(p); }
  static __device__ __forceinline__ void st(float* p, const float (&v)[1]) { *p = v[0]; }
};
template <> struct VecLoad<double, 2> {
  static __device__ __forceinline__ void ld(const double* p, double (&v)[2]) {
    const double2 r = ldg_stream(reinterpret_cast<const double2*>(p));
    v[0] = r.x; v[1] = r.y;
  }
  static __device__ __forceinline__ void st(double* p, const double (&v)[2]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
  }
};
template <> struct VecLoad<double, 1> {
  static __device__ __forceinline__ void ld(const double* p, double (&v)[1]) { v[0] = ldg_stream(p); }
  static __device__ __forceinline__ void st(double* p, const double (&v)[1]) { *p = v[0]; }
};

template <typename T, int VEC, int STAT, bool SKIPNA>
__global__ void __launch_bounds__(kMapThreads) det_maps_kernel(const MapParams p) {
  const int64_t j = blockIdx.x / p.bpm;
  const int64_t vi = int64_t(blockIdx.x % p.bpm) * kMapThreads + threadIdx.x;
  if (vi >= int64_t(p.nrow) * p.ncv) return;
  const int row = static_cast<int>(vi / p.ncv);
  const int col = static_cast<int>(vi % p.ncv) * VEC;
  const int64_t cell = int64_t(row) * p.row_stride + col;
  const T* __restrict__ pf = static_cast<const T*>(p.f) + cell;
  const T* __restrict__ pt = static_cast<const T*>(p.t) + cell;
  const int64_t* __restrict__ of = p.off_f + j * p.ngroup;
  const int64_t* __restrict__ ot = p.off_t + j * p.ngroup;

  double acc[VEC];
  int cnt[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) { acc[e] = 0.0; cnt[e] = 0; }
  T last[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) last[e] = T(0);
#pragma unroll 4
  for (int g = 0; g < p.ngroup; ++g) {
    T a[VEC], b[VEC];
    VecLoad<T, VEC>::ld(pf + of[g], a);
    VecLoad<T, VEC>::ld(pt + ot[g], b);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const T d = a[e] - b[e];
      const T v = STAT == WB2_MAP_BIAS ? d : (STAT == WB2_MAP_MSE ? d * d : (d < T(0) ? -d : d));
      last[e] = v;
      if (SKIPNA) {
        if (v == v) { acc[e] += double(v); ++cnt[e]; }
      } else {
        acc[e] += double(v);
      }
    }
  }
  T res[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    if (p.ngroup == 1) res[e] = last[e];  // compute_chunk: the map itself
    else if (SKIPNA) res[e] = cnt[e] > 0 ? T(acc[e] / double(cnt[e])) : T(nan(""));
    else res[e] = T(acc[e] / double(p.ngroup));
  }
  T* po = static_cast<T*>(p.out) + j * p.cells_per_map + int64_t(row) * p.ncol + col;
  VecLoad<T, VEC>::st(po, res);
}

template <typename T, int VEC>
static int launch_maps(wb2_ctx* ctx, const MapParams& p, int64_t nout, int stat, bool skipna) {
  const dim3 grid(static_cast<unsigned>(nout * p.bpm));
#define WB2_MAP_GO(S)                                                                  \
  (skipna ? det_maps_kernel<T, VEC, S, true><<<grid, kMapThreads, 0, ctx->stream>>>(p) \
          : det_maps_kernel<T, VEC, S, false><<<grid, kMapThreads, 0, ctx->stream>>>(p))
  if (stat == WB2_MAP_BIAS) WB2_MAP_GO(WB2_MAP_BIAS);
  else if (stat == WB2_MAP_MSE) WB2_MAP_GO(WB2_MAP_MSE);
  else WB2_MAP_GO(WB2_MAP_MAE);
#undef WB2_MAP_GO
  WB2_CUDA_TRY(cudaGetLastError());
  return WB2_OK;
}

}  // namespace wb2

using namespace wb2;

extern "C" int wb2_det_maps(wb2_ctx* ctx, const void* f, const void* t, int dtype, int stat,
                            int64_t nout, int32_t ngroup, const int64_t* off_f,
                            const int64_t* off_t, int32_t nrow, int32_t ncol, int64_t row_stride,
                            int skipna, void* out) {
  WB2_NVTX("wb2_det_maps");
  WB2_REQUIRE(ctx != nullptr, "ctx is NULL");
  WB2_REQUIRE(dtype == WB2_F32 || dtype == WB2_F64, "dtype must be WB2_F32 or WB2_F64");
  WB2_REQUIRE(stat == WB2_MAP_BIAS || stat == WB2_MAP_MSE || stat == WB2_MAP_MAE,
              "wb2_det_maps: stat must be WB2_MAP_BIAS, WB2_MAP_MSE or WB2_MAP_MAE");
  WB2_REQUIRE(nrow > 0 && ncol > 0 && row_stride >= ncol, "bad grid: nrow=%d ncol=%d", nrow, ncol);
  WB2_REQUIRE(nout >= 0 && ngroup >= 1, "nout must be >= 0 and ngroup >= 1");
  if (nout == 0) return WB2_OK;
  WB2_REQUIRE(f && t && off_f && off_t && out, "f/t/out and the offset tables must not be NULL");
  DeviceGuard guard(ctx->device);
  const int64_t nfield = nout * ngroup;
  const size_t es = dtype == WB2_F32 ? 4 : 8;
  const int vec = dtype == WB2_F32 ? 4 : 2;
  // 128-bit path: every slab, row and the output start on a 16-byte boundary
  bool aligned = ncol % vec == 0 && row_stride % vec == 0 &&
                 reinterpret_cast<uintptr_t>(f) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(t) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(out) % 16 == 0;
  for (int64_t i = 0; aligned && i < nfield; ++i)
    aligned = off_f[i] % vec == 0 && off_t[i] % vec == 0;
  (void)es;
  Packer pk(ctx);
  const size_t o_f = pk.add(off_f, nfield * sizeof(int64_t));
  const size_t o_t = pk.add(off_t, nfield * sizeof(int64_t));
  WB2_TRY(pk.commit());
  MapParams p;
  p.f = f; p.t = t; p.out = out;
  p.off_f = pk.dev<int64_t>(o_f);
  p.off_t = pk.dev<int64_t>(o_t);
  p.row_stride = row_stride;
  p.cells_per_map = int64_t(nrow) * ncol;
  p.ngroup = ngroup; p.nrow = nrow; p.ncol = ncol;
  const int v = aligned ? vec : 1;
  p.ncv = ncol / v;
  const int64_t nv = int64_t(nrow) * p.ncv;
  p.bpm = static_cast<int32_t>((nv + kMapThreads - 1) / kMapThreads);
  WB2_REQUIRE(nout * int64_t(p.bpm) < (int64_t(1) << 31), "launch too large");
  int rc;
  if (dtype == WB2_F32)
    rc = aligned ? launch_maps<float, 4>(ctx, p, nout, stat, skipna != 0)
                 : launch_maps<float, 1>(ctx, p, nout, stat, skipna != 0);
  else
    rc = aligned ? launch_maps<double, 2>(ctx, p, nout, stat, skipna != 0)
                 : launch_maps<double, 1>(ctx, p, nout, stat, skipna != 0);
  if (rc != WB2_OK) return rc;
  ctx->launches += 1;
  WB2_TRY(pk.release());
  return WB2_OK;
}
