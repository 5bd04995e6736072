// wb2_det_metrics_host -- end-to-end entry for HOST-resident inputs.
//
// The reference calls Metric.compute_chunk on NumPy-backed xarray chunks
// (weatherbench2/evaluation.py:583-599); this entry takes the same host
// buffers, streams the 2-D slabs through two device staging buffers (H2D on a
// copy stream, overlapped with the K1 kernel of the previous group) and
// returns the float64 sums in host memory.  PCIe, not HBM, bounds this path.
#include <unordered_map>

#include "common.cuh"

namespace wb2 {

int det_metrics_impl(wb2_ctx* ctx, int mode, const void* f, const void* t, const void* c,
                     const void* g, int dtype, int64_t nfield, const int64_t* off_f,
                     const int64_t* off_t, const int64_t* off_c, const int64_t* off_g,
                     const wb2_weights* w, int skipna, double* out);

static size_t stage_bytes_default() {
  const char* env = getenv("WB2_STAGE_MB");
  size_t mb = env ? strtoull(env, nullptr, 10) : 128;
  if (mb < 1) mb = 1;
  return mb << 20;
}

static int ensure_stage(wb2_ctx* ctx, size_t need) {
  if (ctx->stage_cap >= need) return WB2_OK;
  WB2_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  WB2_CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
  for (int i = 0; i < 2; ++i) {
    if (ctx->stage[i]) WB2_CUDA_TRY(cudaFree(ctx->stage[i]));
    ctx->stage[i] = nullptr;
  }
  ctx->stage_cap = 0;
  for (int i = 0; i < 2; ++i) WB2_CUDA_TRY(cudaMalloc(&ctx->stage[i], need));
  ctx->stage_cap = need;
  return WB2_OK;
}

static int ensure_out_tmp(wb2_ctx* ctx, size_t need) {
  if (ctx->out_tmp_cap >= need) return WB2_OK;
  WB2_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (ctx->d_out_tmp) WB2_CUDA_TRY(cudaFree(ctx->d_out_tmp));
  ctx->d_out_tmp = nullptr;
  ctx->out_tmp_cap = 0;
  WB2_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&ctx->d_out_tmp), need));
  ctx->out_tmp_cap = need;
  return WB2_OK;
}

}  // namespace wb2

using namespace wb2;

extern "C" int wb2_det_metrics_host(wb2_ctx* ctx, const void* f, const void* t, const void* c,
                                    int dtype, int64_t nfield, const int64_t* off_f,
                                    const int64_t* off_t, const int64_t* off_c,
                                    const wb2_weights* w, int skipna, double* out_host) {
  WB2_REQUIRE(ctx != nullptr, "ctx is NULL");
  WB2_REQUIRE(dtype == WB2_F32 || dtype == WB2_F64, "dtype must be WB2_F32 or WB2_F64");
  WB2_TRY(validate_weights(w));
  WB2_REQUIRE(out_host != nullptr, "out is NULL");
  if (nfield == 0) return WB2_OK;
  WB2_REQUIRE(f && t && off_f && off_t, "f/t and their offset tables must not be NULL");
  if (c) WB2_REQUIRE(off_c != nullptr, "climatology offsets are NULL");
  DeviceGuard guard(ctx->device);

  const size_t esize = dtype == WB2_F32 ? 4 : 8;
  const int noper = c ? 3 : 2;
  const int64_t slab_elems = int64_t(w->nrow - 1) * w->row_stride + w->ncol;
  const int64_t slab_pad = (slab_elems + 63) / 64 * 64;  // keeps 256-B alignment
  const size_t slab_bytes = size_t(slab_pad) * esize;
  size_t cap = stage_bytes_default();
  if (cap < slab_bytes * noper) cap = slab_bytes * noper;
  WB2_TRY(ensure_stage(ctx, cap));
  const int64_t max_slabs = int64_t(ctx->stage_cap / slab_bytes);
  const size_t per_field = size_t(w->nregion) * WB2_DET_NSTAT;
  WB2_TRY(ensure_out_tmp(ctx, size_t(nfield) * per_field * sizeof(double)));

  const char* hosts[3] = {static_cast<const char*>(f), static_cast<const char*>(t),
                          static_cast<const char*>(c)};
  const int64_t* offs[3] = {off_f, off_t, off_c};

  std::vector<int64_t> loc[3];
  int buf = 0;
  int64_t g0 = 0;
  bool used[2] = {false, false};
  while (g0 < nfield) {
    // greedily take fields while their (deduplicated) slabs fit the buffer
    std::unordered_map<const char*, int64_t> slot_of;  // host address -> slab slot
    for (int o = 0; o < noper; ++o) loc[o].clear();
    int64_t g1 = g0;
    while (g1 < nfield) {
      int fresh = 0;
      for (int o = 0; o < noper; ++o)
        if (!slot_of.count(hosts[o] + offs[o][g1] * esize)) ++fresh;
      if (int64_t(slot_of.size()) + fresh > max_slabs) break;
      for (int o = 0; o < noper; ++o) {
        const char* src = hosts[o] + offs[o][g1] * esize;
        auto it = slot_of.find(src);
        if (it == slot_of.end()) it = slot_of.emplace(src, int64_t(slot_of.size())).first;
        loc[o].push_back(it->second * slab_pad);
      }
      ++g1;
    }
    WB2_REQUIRE(g1 > g0, "staging buffer too small for one field");
    char* stage = static_cast<char*>(ctx->stage[buf]);
    if (used[buf]) WB2_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->stage_free[buf], 0));
    for (const auto& kv : slot_of) {
      WB2_CUDA_TRY(cudaMemcpyAsync(stage + size_t(kv.second) * slab_bytes, kv.first,
                                   size_t(slab_elems) * esize, cudaMemcpyHostToDevice,
                                   ctx->copy_stream));
    }
    WB2_CUDA_TRY(cudaEventRecord(ctx->stage_copied[buf], ctx->copy_stream));
    WB2_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->stage_copied[buf], 0));
    int rc = det_metrics_impl(ctx, c ? 1 : 0, stage, stage, c ? stage : nullptr, nullptr, dtype,
                              g1 - g0, loc[0].data(), loc[1].data(),
                              c ? loc[2].data() : nullptr, nullptr, w, skipna,
                              ctx->d_out_tmp + size_t(g0) * per_field);
    if (rc != WB2_OK) return rc;
    WB2_CUDA_TRY(cudaEventRecord(ctx->stage_free[buf], ctx->stream));
    used[buf] = true;
    buf ^= 1;
    g0 = g1;
  }
  WB2_CUDA_TRY(cudaMemcpyAsync(out_host, ctx->d_out_tmp, size_t(nfield) * per_field * sizeof(double),
                               cudaMemcpyDeviceToHost, ctx->stream));
  WB2_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return WB2_OK;
}
