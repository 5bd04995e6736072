// K10 -- rank histogram of the truth within the ensemble (sm_100a).
//
// Replaces RankHistogram.compute_chunk (weatherbench2/metrics.py:1894-2042) and,
// for ngroup > 1, the time mean of EnsembleMetric.compute.  The reference
// concatenates truth and members, optionally perturbs every value by uniform
// noise smaller than a quarter of the smallest positive gap (:1960-1987),
// argsorts, and one-hot encodes the position of the truth (:2019-2037).
// The noise cannot reorder distinct values, so
//     rank = #{members < truth} + J,   J uniform on {0, ..., k},
// with k the number of members exactly equal to the truth (J = 0 without
// random tie-breaking: the truth is first in the concatenation and a stable
// sort keeps it first).  NaN sorts last: a NaN truth ranks after every valid
// member, NaN members are never below the truth.  Members are streamed once
// (4 M + 4 bytes per grid point); the only random draw happens on exact ties
// and comes from a counter-based hash of (seed, time slab, cell), so results
// are reproducible for a given seed but are not NumPy's PCG64 stream -- on
// ties the parity with the reference is distributional, everywhere else exact.
#include "common.cuh"

namespace wb2 {

constexpr int kRankThreads = 128;

struct RankParams {
  const float* x;
  const float* t;
  float* out;             // [nout][nrow][ncol][nbins]
  const int64_t* off_x;   // [nout][ngroup]
  const int64_t* off_t;
  int64_t member_stride, row_stride, cells_per_map;
  uint64_t seed;
  int32_t nmember, ngroup, nrow, ncol, nbins, reduction, bpm, random_ties;
};

__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(kRankThreads) rank_hist_kernel(const RankParams p) {
  const int64_t j = blockIdx.x / p.bpm;
  const int64_t ci = int64_t(blockIdx.x % p.bpm) * kRankThreads + threadIdx.x;
  if (ci >= p.cells_per_map) return;
  const int row = static_cast<int>(ci / p.ncol);
  const int col = static_cast<int>(ci % p.ncol);
  const int64_t cell = int64_t(row) * p.row_stride + col;
  float* o = p.out + (j * p.cells_per_map + ci) * p.nbins;
  for (int b = 0; b < p.nbins; ++b) o[b] = 0.f;
  for (int g = 0; g < p.ngroup; ++g) {
    const int64_t field = j * p.ngroup + g;
    const float t = ldg_stream(p.t + p.off_t[field] + cell);
    const float* src = p.x + p.off_x[field] + cell;
    int below = 0, equal = 0, valid = 0;
#pragma unroll 10
    for (int m = 0; m < p.nmember; ++m) {
      const float xm = ldg_stream(src + int64_t(m) * p.member_stride);
      below += xm < t ? 1 : 0;
      equal += xm == t ? 1 : 0;
      valid += xm == xm ? 1 : 0;
    }
    int rank;
    if (!(t == t)) {
      rank = valid;  // NaN sorts last; the truth comes first among the NaNs
    } else {
      rank = below;
      if (p.random_ties && equal > 0) {
        const uint64_t h = splitmix64(splitmix64(p.seed ^ uint64_t(field)) ^ uint64_t(ci));
        // uniform on {0, ..., equal}: high bits of a 64-bit hash
        rank += static_cast<int>(__umul64hi(h, uint64_t(equal + 1)));
      }
    }
    const int bin = rank / p.reduction;  // metrics.py:1950-1958
    // one thread owns the cell: no atomics; integer counts are exact in float32
    o[bin] += 1.f;
  }
  if (p.ngroup > 1)
    for (int b = 0; b < p.nbins; ++b) o[b] = o[b] / float(p.ngroup);  // mean of the one-hots
}

}  // namespace wb2

using namespace wb2;

extern "C" int wb2_rank_histogram(wb2_ctx* ctx, const float* x, const float* t,
                                  int32_t nmember, int64_t member_stride, int64_t nout,
                                  int32_t ngroup, const int64_t* off_x, const int64_t* off_t,
                                  int32_t nrow, int32_t ncol, int64_t row_stride, int32_t nbins,
                                  int32_t random_ties, uint64_t seed, float* out) {
  WB2_NVTX("wb2_rank_histogram");
  WB2_REQUIRE(ctx != nullptr, "ctx is NULL");
  WB2_REQUIRE(nmember >= 1 && nmember <= 65536, "nmember out of range");
  WB2_REQUIRE(nbins >= 1 && (nmember + 1) % nbins == 0,
              "Cannot bin data with ensemble_size=%d into %d bins", nmember, nbins);
  WB2_REQUIRE(nrow > 0 && ncol > 0 && row_stride >= ncol, "bad grid: nrow=%d ncol=%d", nrow, ncol);
  WB2_REQUIRE(nout >= 0 && ngroup >= 1, "nout must be >= 0 and ngroup >= 1");
  if (nout == 0) return WB2_OK;
  WB2_REQUIRE(x && t && off_x && off_t && out, "NULL argument");
  DeviceGuard guard(ctx->device);
  const int64_t nfield = nout * ngroup;
  Packer pk(ctx);
  const size_t o_x = pk.add(off_x, nfield * sizeof(int64_t));
  const size_t o_t = pk.add(off_t, nfield * sizeof(int64_t));
  WB2_TRY(pk.commit());
  RankParams p;
  p.x = x; p.t = t; p.out = out;
  p.off_x = pk.dev<int64_t>(o_x);
  p.off_t = pk.dev<int64_t>(o_t);
  p.member_stride = member_stride;
  p.row_stride = row_stride;
  p.cells_per_map = int64_t(nrow) * ncol;
  p.seed = seed;
  p.nmember = nmember; p.ngroup = ngroup; p.nrow = nrow; p.ncol = ncol;
  p.nbins = nbins;
  p.reduction = (nmember + 1) / nbins;
  p.bpm = static_cast<int32_t>((p.cells_per_map + kRankThreads - 1) / kRankThreads);
  p.random_ties = random_ties ? 1 : 0;
  WB2_REQUIRE(nout * int64_t(p.bpm) < (int64_t(1) << 31), "launch too large");
  rank_hist_kernel<<<static_cast<unsigned>(nout * p.bpm), kRankThreads, 0, ctx->stream>>>(p);
  WB2_CUDA_TRY(cudaGetLastError());
  ctx->launches += 1;
  WB2_TRY(pk.release());
  return WB2_OK;
}
