// K6e -- map-output ensemble metrics with the time mean fused in (sm_100a).
//
// Replaces SpatialCRPS / SpatialCRPSSkill / SpatialCRPSSpread
// (weatherbench2/metrics.py:718-772), SpatialEnsembleVariance (:1244-1266),
// SpatialEnsembleMeanMSE / DebiasedSpatialEnsembleMeanMSE (:1366-1399)
// .compute_chunk and, for ngroup > 1, the `.mean(time, skipna)` of
// EnsembleMetric.compute (:598-607) in the same pass.  The point-wise math is
// K2's (ens_point.cuh: members in registers, Batcher sorting network); every
// selected statistic is one float32 map per output.  One read of the M
// members yields all selected maps: (4 M + 4) bytes per grid point per time
// step + 4 * nsel / ngroup written.
#include "common.cuh"
#include "ens_point.cuh"

namespace wb2 {

constexpr int kEnsMapThreads = 128;
constexpr int kMapStats = kEnsStats + 1;  // + point-wise CRPS = skill - spread / 2

struct EnsMapParams {
  const float* x;
  const float* t;
  float* out;  // [nsel][nout][nrow][ncol]
  const int64_t* off_x;  // [nout][ngroup]
  const int64_t* off_t;
  int64_t member_stride, row_stride, cells_per_map, sel_stride;
  int32_t nmember, ngroup, nrow, ncol, bpm, stat_mask;
};

template <int MP, bool SKIPNA, bool EXACT>
__global__ void __launch_bounds__(kEnsMapThreads, 3) ens_maps_kernel(const EnsMapParams p) {
  const int64_t j = blockIdx.x / p.bpm;
  const int64_t ci = int64_t(blockIdx.x % p.bpm) * kEnsMapThreads + threadIdx.x;
  if (ci >= p.cells_per_map) return;
  const int row = static_cast<int>(ci / p.ncol);
  const int col = static_cast<int>(ci % p.ncol);
  const int64_t cell = int64_t(row) * p.row_stride + col;
  const int M = p.nmember;
  double acc[kMapStats];
  int cnt[kMapStats];
  float last[kMapStats];
#pragma unroll
  for (int i = 0; i < kMapStats; ++i) { acc[i] = 0.0; cnt[i] = 0; last[i] = 0.f; }
  for (int g = 0; g < p.ngroup; ++g) {
    const float* __restrict__ src = p.x + p.off_x[j * p.ngroup + g] + cell;
    float v[MP];
#pragma unroll
    for (int m = 0; m < MP; ++m) {
      if (EXACT || m < M) v[m] = ldg_stream(src + int64_t(m) * p.member_stride);
      else v[m] = 0.f;
    }
    const float t = ldg_stream(p.t + p.off_t[j * p.ngroup + g] + cell);
    float pt[kEnsStats];
    ens_point<MP, SKIPNA, EXACT, kTwinSortOk<MP, SKIPNA, EXACT>>(v, t, M, pt);
    float val[kMapStats];
#pragma unroll
    for (int i = 0; i < kEnsStats; ++i) val[i] = pt[i];
    val[5] = pt[0] - 0.5f * pt[1];  // point-wise CRPS (metrics.py:729-739)
#pragma unroll
    for (int i = 0; i < kMapStats; ++i) {
      last[i] = val[i];
      if (SKIPNA) {
        if (val[i] == val[i]) { acc[i] += double(val[i]); ++cnt[i]; }
      } else {
        acc[i] += double(val[i]);
      }
    }
  }
  float* po = p.out + j * p.cells_per_map + ci;
#pragma unroll
  for (int i = 0; i < kMapStats; ++i) {
    if (!(p.stat_mask & (1 << i))) continue;
    float r;
    if (p.ngroup == 1) r = last[i];
    else if (SKIPNA) r = cnt[i] > 0 ? float(acc[i] / double(cnt[i])) : __int_as_float(0x7fc00000);
    else r = float(acc[i] / double(p.ngroup));
    *po = r;
    po += p.sel_stride;
  }
}

template <int MP>
static int launch_ens_maps(wb2_ctx* ctx, const EnsMapParams& p, int64_t nout, bool skipna) {
  const bool exact = p.nmember == MP;
  const dim3 grid(static_cast<unsigned>(nout * p.bpm));
  if (skipna) {
    if (exact) ens_maps_kernel<MP, true, true><<<grid, kEnsMapThreads, 0, ctx->stream>>>(p);
    else ens_maps_kernel<MP, true, false><<<grid, kEnsMapThreads, 0, ctx->stream>>>(p);
  } else {
    if (exact) ens_maps_kernel<MP, false, true><<<grid, kEnsMapThreads, 0, ctx->stream>>>(p);
    else ens_maps_kernel<MP, false, false><<<grid, kEnsMapThreads, 0, ctx->stream>>>(p);
  }
  WB2_CUDA_TRY(cudaGetLastError());
  return WB2_OK;
}

}  // namespace wb2

using namespace wb2;

extern "C" int wb2_ens_maps(wb2_ctx* ctx, const void* x, const void* t, int dtype,
                            int32_t nmember, int64_t member_stride, int64_t nout, int32_t ngroup,
                            const int64_t* off_x, const int64_t* off_t, int32_t nrow,
                            int32_t ncol, int64_t row_stride, int32_t stat_mask, int skipna,
                            float* out) {
  WB2_NVTX("wb2_ens_maps");
  WB2_REQUIRE(ctx != nullptr, "ctx is NULL");
  WB2_REQUIRE(dtype == WB2_F32, "wb2_ens_maps: only WB2_F32 inputs are supported");
  if (nmember < 1 || nmember > 64) {
    set_error("wb2_ens_maps: 1..64 ensemble members are supported (got %d)", nmember);
    return nmember < 1 ? WB2_EINVAL : WB2_EUNSUPPORTED;
  }
  WB2_REQUIRE(stat_mask > 0 && stat_mask < (1 << kMapStats), "stat_mask must select 1..6 maps");
  WB2_REQUIRE(nrow > 0 && ncol > 0 && row_stride >= ncol, "bad grid: nrow=%d ncol=%d", nrow, ncol);
  WB2_REQUIRE(nout >= 0 && ngroup >= 1, "nout must be >= 0 and ngroup >= 1");
  if (nout == 0) return WB2_OK;
  WB2_REQUIRE(x && t && off_x && off_t && out, "x/t/out and the offset tables must not be NULL");
  DeviceGuard guard(ctx->device);
  const int64_t nfield = nout * ngroup;
  Packer pk(ctx);
  const size_t o_x = pk.add(off_x, nfield * sizeof(int64_t));
  const size_t o_t = pk.add(off_t, nfield * sizeof(int64_t));
  WB2_TRY(pk.commit());
  EnsMapParams p;
  p.x = static_cast<const float*>(x);
  p.t = static_cast<const float*>(t);
  p.out = out;
  p.off_x = pk.dev<int64_t>(o_x);
  p.off_t = pk.dev<int64_t>(o_t);
  p.member_stride = member_stride;
  p.row_stride = row_stride;
  p.cells_per_map = int64_t(nrow) * ncol;
  p.sel_stride = nout * p.cells_per_map;
  p.nmember = nmember; p.ngroup = ngroup; p.nrow = nrow; p.ncol = ncol;
  p.bpm = static_cast<int32_t>((p.cells_per_map + kEnsMapThreads - 1) / kEnsMapThreads);
  p.stat_mask = stat_mask;
  WB2_REQUIRE(nout * int64_t(p.bpm) < (int64_t(1) << 31), "launch too large");
  const bool sk = skipna != 0;
  int rc;
  if (nmember <= 2) rc = launch_ens_maps<2>(ctx, p, nout, sk);
  else if (nmember <= 3) rc = launch_ens_maps<3>(ctx, p, nout, sk);
  else if (nmember <= 4) rc = launch_ens_maps<4>(ctx, p, nout, sk);
  else if (nmember <= 5) rc = launch_ens_maps<5>(ctx, p, nout, sk);
  else if (nmember <= 8) rc = launch_ens_maps<8>(ctx, p, nout, sk);
  else if (nmember <= 10) rc = launch_ens_maps<10>(ctx, p, nout, sk);
  else if (nmember <= 16) rc = launch_ens_maps<16>(ctx, p, nout, sk);
  else if (nmember <= 20) rc = launch_ens_maps<20>(ctx, p, nout, sk);
  else if (nmember <= 32) rc = launch_ens_maps<32>(ctx, p, nout, sk);
  else if (nmember <= 50) rc = launch_ens_maps<50>(ctx, p, nout, sk);
  else rc = launch_ens_maps<64>(ctx, p, nout, sk);
  if (rc != WB2_OK) return rc;
  ctx->launches += 1;
  WB2_TRY(pk.release());
  return WB2_OK;
}
