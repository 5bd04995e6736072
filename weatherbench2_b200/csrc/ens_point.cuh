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
// Twin compare-exchange on two aligned register pairs (v[a], v[a+1]) and
// (v[b], v[b+1]): the minima still cost one FMNMX each (ALU pipe, one warp
// instruction per 2 cycles per SM sub-partition), but the maxima come from
// hi = (a + b) - lo as two packed FADD2 on the FMA pipe, which halves the ALU
// work of those comparators.  (a + b) - lo is not exact: it is only used on
// mean-removed members, where the rounding error is relative to the ensemble
// spread (see ens_point below and DESIGN.md, K2).
__device__ __forceinline__ void ce2_packed(float& a0, float& a1, float& b0, float& b1) {
  const float l0 = fminf(a0, b0);
  const float l1 = fminf(a1, b1);
  unsigned long long pa, pb, pl, ps, ph;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pl) : "f"(l0), "f"(l1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(ps) : "l"(pa), "l"(pb));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(ph) : "l"(ps), "l"(pl));
  a0 = l0;
  a1 = l1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(b0), "=f"(b1) : "l"(ph));
}
#define CE2(a, b) ce2_packed(v[a], v[(a) + 1], v[b], v[(b) + 1]);
#include "sort_networks.inc"
#undef CE
#undef CE2

constexpr int kEnsStats = 5;

// The packed network applies to full (unpadded), even-sized ensembles without
// NaN handling: padding is +inf and inf + inf - inf is NaN.
template <int MP, bool SKIPNA, bool EXACT>
constexpr bool kTwinSortOk = EXACT && !SKIPNA && MP >= 8 && MP % 2 == 0;

// Point-wise statistics of one grid point.  x[m] for m >= M is padding.
//   [0] skill_pt  = mean_m |t - x_m|                          (metrics.py:824)
//   [1] spread_pt = 2 * mean_m((2 r_m - M - 1) x_m) / (M - 1) (metrics.py:805-813)
//   [2] (t - xbar)^2          [3] var_m(x, ddof=1)     [4] [2] - [3] / M
// TWIN: sort the mean-removed members with the packed network (the spread sum
// is invariant to a common shift because its coefficients add up to zero).
template <int MP, bool SKIPNA, bool EXACT, bool TWIN = false>
__device__ __forceinline__ void ens_point(float (&v)[MP], float t, int M, float (&val)[kEnsStats]) {
  static_assert(!TWIN || kTwinSortOk<MP, SKIPNA, EXACT>, "packed sort needs a full even ensemble");
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
      if (TWIN) v[m] = dx;
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
  if constexpr (TWIN) SortNetTwin<MP>::run(v);
  else SortNet<MP>::run(v);
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
