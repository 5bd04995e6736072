// Point-wise ensemble statistics shared by K2 (ens_metrics.cu, spatially
// averaged) and K6e (ens_maps.cu, map output): the M members of one grid point
// sit in registers and pass through a Batcher sorting network.
#pragma once
#include "common.cuh"

namespace wb2 {

#define CE(a, b)                          \
  {                                       \
    const float lo_ = fminf(v[a], v[b]);  \
    const float hi_ = fmaxf(v[a], v[b]);  \
    v[a] = lo_;                           \
    v[b] = hi_;                           \
  }
#include "sort_networks.inc"
#undef CE

constexpr int kEnsStats = 5;

// Point-wise statistics of one grid point.  x[m] for m >= M is padding.
//   [0] skill_pt  = mean_m |t - x_m|                          (metrics.py:824)
//   [1] spread_pt = 2 * mean_m((2 r_m - M - 1) x_m) / (M - 1) (metrics.py:805-813)
//   [2] (t - xbar)^2          [3] var_m(x, ddof=1)     [4] [2] - [3] / M
template <int MP, bool SKIPNA, bool EXACT>
__device__ __forceinline__ void ens_point(float (&v)[MP], float t, int M, float (&val)[kEnsStats]) {
  const float nanf_ = __int_as_float(0x7fc00000);
  const float inf_ = __int_as_float(0x7f800000);
  float sumx = 0.f, suma = 0.f;
  float nvalid = 0.f, navalid = 0.f;
#pragma unroll
  for (int m = 0; m < MP; ++m) {
    if (EXACT || m < M) {
      const float xm = v[m];
      const float a = fabsf(t - xm);  // metrics.py:824
      if (SKIPNA) {
        if (xm == xm) { sumx += xm; nvalid += 1.f; }
        if (a == a) { suma += a; navalid += 1.f; }
      } else {
        sumx += xm;
        suma += a;
      }
    }
  }
  const float fm = float(M);
  const float mean = SKIPNA ? sumx / nvalid : sumx / fm;  // 0/0 -> NaN like nanmean
  float ss = 0.f;
#pragma unroll
  for (int m = 0; m < MP; ++m) {
    if (EXACT || m < M) {
      const float dx = v[m] - mean;
      if (SKIPNA) {
        if (dx == dx) ss += dx * dx;
      } else {
        ss += dx * dx;
      }
    }
  }
  float var;
  if (SKIPNA) var = nvalid > 1.f ? ss / (nvalid - 1.f) : nanf_;  // np.nanvar(ddof=1)
  else var = ss / (fm - 1.f);                                    // M == 1 -> 0/0 = NaN
  const float dm = t - mean;
  const float mse = dm * dm;
  val[0] = SKIPNA ? suma / navalid : suma / fm;
  val[2] = mse;
  val[3] = var;
  val[4] = mse - var / fm;  // metrics.py:564-565 (always divides by the full M)

  // ---- spread: ranks via sorting network (metrics.py:804-813) ---------------
  if (M < 2) {
    val[1] = 0.f;  // metrics.py:788-789
    return;
  }
#pragma unroll
  for (int m = 0; m < MP; ++m) {
    if (!EXACT && m >= M) v[m] = inf_;          // padding sorts last
    else if (SKIPNA && !(v[m] == v[m])) v[m] = inf_;  // NaN sorts last (np.argsort)
  }
  SortNet<MP>::run(v);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MP; ++i) {
    // coefficient 2 r - M - 1 with r = i + 1
    const float coef = float(2 * (i + 1)) - fm - 1.f;
    if (SKIPNA) {
      if (float(i) < nvalid) s += coef * v[i];
    } else if (EXACT || i < M) {
      s += coef * v[i];
    }
  }
  float spread = SKIPNA ? 2.f * (s / nvalid) / (fm - 1.f) : 2.f * (s / fm) / (fm - 1.f);
  if (!SKIPNA && !(sumx == sumx)) spread = nanf_;  // a NaN member poisons the point
  val[1] = spread;
}

}  // namespace wb2
