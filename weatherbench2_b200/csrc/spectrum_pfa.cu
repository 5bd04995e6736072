// K4 fast path -- zonal energy spectrum as a prime-factor FFT in packed
// f32x2 arithmetic, rows staged by TMA (sm_100a).
//
// Replaces ZonalEnergySpectrum.compute (weatherbench2/derived_variables.py:
// 592-626) for the row lengths whose half N2 = N / 2 splits into three pairwise
// coprime radices RA * RB * RC (1440 longitudes: N2 = 720 = 9 * 16 * 5), plus
// the reductions that follow it in the callers: the time mean of
// scripts/compute_zonal_energy_spectrum.py:234 and (north star) the
// latitude-weighted meridional reduction with get_lat_weights
// (weatherbench2/metrics.py:40-60).
//
// Why this shape (round-1 K4 ran at 0.27 of the HBM roofline, FMA-pipe bound:
// 618 scalar FP + 207 IMAD instructions per butterfly loop on a pipe that
// issues one 3-register FFMA per 2 cycles and SMSP):
//   * PACKED ARITHMETIC.  Two latitude rows (A, B) are transformed together:
//     every complex value is held as two 64-bit register pairs
//     re = (re_A, re_B), im = (im_A, im_B) and every operation is an
//     add/sub/mul/fma.rn.f32x2 (SASS FADD2 / FMUL2 / FFMA2) -- half the FP
//     instructions per row, and because the two lanes are two independent rows
//     there is not a single swizzle: multiplying by -i is a register rename.
//   * PRIME-FACTOR (Good-Thomas) decomposition: with the input gathered by
//     n = (SA nA + SB nB + SC nC) mod N2 (S* = N2 / R*) and the output read at
//     k = (EA kA + EB kB + EC kC) mod N2 (E* the CRT idempotents), the N2-point
//     DFT is three plain DFTs of size RA, RB, RC along the axes of a
//     [RB][RA][RC] array: NO inter-stage twiddles (the Stockham version loaded
//     and multiplied 1 216 of them per row), IN PLACE (one shared buffer, no
//     ping-pong, no padding), and every index is a compile-time stride.
//   * The last stage computes butterfly (kA, kB) TOGETHER with its mirror
//     (-kA, -kB): bin p and bin N2 - p of the real-input split
//         X_p = (E + W_N^p O) / 2,  X_{N2-p} = conj(E - W_N^p O) / 2
//     are then both in the same thread's registers, so the power spectrum is
//     accumulated (over time steps, or over latitude rows with their weights)
//     straight from registers; the spectrum itself never goes to shared memory
//     unless it has to be written.
//   * TMA: the 2 G rows of the next work item are fetched by one elected thread
//     with cp.async.bulk (SASS UBLKCP) into a staging buffer guarded by an
//     mbarrier while stages B and C of the current item run.
//
// Work decomposition: a work item = G row pairs (2 G rows) of one field; a job
// = the items that share accumulators (mode 0: one item; mode 1: the same rows
// of every time step of a slot; mode 2: a chunk of row groups of every time
// step of a slot).  Persistent CTAs (2 per SM) take jobs round-robin; the
// accumulators are flushed through shared memory (coalesced stores) at the end
// of a job.  Deterministic: no atomics, fixed summation order.
//
// Shared-memory traffic per row: staging read 45 + three in-place passes
// (45 * 4) + post twiddles ~ 20 cycles at 128 B/cycle against 246 cycles per
// row at the HBM roofline; FMA pipe ~ 430 packed instructions per row.
#include <cmath>

#include "common.cuh"
#include "tma_utils.cuh"

namespace wb2 {
namespace pfa {

typedef unsigned long long u64;

// two complex numbers (lane 0 = row A, lane 1 = row B), split re / im
struct C2 {
  u64 re, im;
};

__device__ __forceinline__ u64 pk2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
// same, but opaque to the optimiser: one pack per value (ptxas otherwise
// re-materialises the two MOVs at every use)
__device__ __forceinline__ u64 pk2v(float lo, float hi) {
  u64 r;
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ u64 dup2(float c) { return pk2(c, c); }
__device__ __forceinline__ float lo2(u64 v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo;
}
__device__ __forceinline__ float hi2(u64 v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return hi;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
  u64 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ C2 cadd(C2 a, C2 b) { return {add2(a.re, b.re), add2(a.im, b.im)}; }
__device__ __forceinline__ C2 csub(C2 a, C2 b) { return {sub2(a.re, b.re), sub2(a.im, b.im)}; }
// a * (wr + i wi), wr / wi compile-time constants (they become FFMA2 immediates)
__device__ __forceinline__ C2 cmulc(C2 a, float wr, float wi) {
  C2 r;
  r.re = fma2(a.im, dup2(-wi), mul2(a.re, dup2(wr)));
  r.im = fma2(a.im, dup2(wr), mul2(a.re, dup2(wi)));
  return r;
}

template <int R>
__device__ __forceinline__ void dft(C2 (&v)[R]);

template <>
__device__ __forceinline__ void dft<1>(C2 (&)[1]) {}
template <>
__device__ __forceinline__ void dft<2>(C2 (&v)[2]) {
  const C2 a = v[0], b = v[1];
  v[0] = cadd(a, b);
  v[1] = csub(a, b);
}
template <>
__device__ __forceinline__ void dft<3>(C2 (&v)[3]) {
  constexpr float S = 0.86602540378443864676f;
  const C2 t = cadd(v[1], v[2]), d = csub(v[1], v[2]);
  const C2 m = {fma2(t.re, dup2(-0.5f), v[0].re), fma2(t.im, dup2(-0.5f), v[0].im)};
  v[0] = cadd(v[0], t);
  // n = -i S d = S (d.im, -d.re);  v1 = m + n, v2 = m - n
  v[1] = {fma2(d.im, dup2(S), m.re), fma2(d.re, dup2(-S), m.im)};
  v[2] = {fma2(d.im, dup2(-S), m.re), fma2(d.re, dup2(S), m.im)};
}
template <>
__device__ __forceinline__ void dft<4>(C2 (&v)[4]) {
  const C2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
  const C2 t2 = cadd(v[1], v[3]), d = csub(v[1], v[3]);
  v[0] = cadd(t0, t2);
  v[2] = csub(t0, t2);
  v[1] = {add2(t1.re, d.im), sub2(t1.im, d.re)};  // t1 - i d
  v[3] = {sub2(t1.re, d.im), add2(t1.im, d.re)};  // t1 + i d
}
template <>
__device__ __forceinline__ void dft<5>(C2 (&v)[5]) {
  constexpr float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
  constexpr float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
  const C2 a = v[0];
  const C2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]);
  const C2 t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
  const C2 m1 = {fma2(t2.re, dup2(c2), fma2(t1.re, dup2(c1), a.re)),
                 fma2(t2.im, dup2(c2), fma2(t1.im, dup2(c1), a.im))};
  const C2 m2 = {fma2(t2.re, dup2(c1), fma2(t1.re, dup2(c2), a.re)),
                 fma2(t2.im, dup2(c1), fma2(t1.im, dup2(c2), a.im))};
  // u1 = s1 t3 + s2 t4, u2 = s2 t3 - s1 t4;  n = -i u = (u.im, -u.re)
  const C2 u1 = {fma2(t4.re, dup2(s2), mul2(t3.re, dup2(s1))),
                 fma2(t4.im, dup2(s2), mul2(t3.im, dup2(s1)))};
  const C2 u2 = {fma2(t4.re, dup2(-s1), mul2(t3.re, dup2(s2))),
                 fma2(t4.im, dup2(-s1), mul2(t3.im, dup2(s2)))};
  v[0] = cadd(a, cadd(t1, t2));
  v[1] = {add2(m1.re, u1.im), sub2(m1.im, u1.re)};
  v[4] = {sub2(m1.re, u1.im), add2(m1.im, u1.re)};
  v[2] = {add2(m2.re, u2.im), sub2(m2.im, u2.re)};
  v[3] = {sub2(m2.re, u2.im), add2(m2.im, u2.re)};
}

// cos / sin of 2 pi m / N for the composite radices (compile-time indices)
template <int N>
__device__ __forceinline__ float wcos(int m);
template <int N>
__device__ __forceinline__ float wsin(int m);
template <>
__device__ __forceinline__ float wcos<16>(int m) {
  constexpr float c[16] = {1.f, 0.92387953251128674f, 0.70710678118654752f,
                           0.38268343236508977f, 0.f, -0.38268343236508977f,
                           -0.70710678118654752f, -0.92387953251128674f, -1.f,
                           -0.92387953251128674f, -0.70710678118654752f,
                           -0.38268343236508977f, 0.f, 0.38268343236508977f,
                           0.70710678118654752f, 0.92387953251128674f};
  return c[m];
}
template <>
__device__ __forceinline__ float wsin<16>(int m) {
  constexpr float s[16] = {0.f, 0.38268343236508977f, 0.70710678118654752f,
                           0.92387953251128674f, 1.f, 0.92387953251128674f,
                           0.70710678118654752f, 0.38268343236508977f, 0.f,
                           -0.38268343236508977f, -0.70710678118654752f,
                           -0.92387953251128674f, -1.f, -0.92387953251128674f,
                           -0.70710678118654752f, -0.38268343236508977f};
  return s[m];
}
template <>
__device__ __forceinline__ float wcos<9>(int m) {
  constexpr float c[9] = {1.f, 0.76604444311897804f, 0.17364817766693035f, -0.5f,
                          -0.93969262078590838f, -0.93969262078590838f, -0.5f,
                          0.17364817766693035f, 0.76604444311897804f};
  return c[m];
}
template <>
__device__ __forceinline__ float wsin<9>(int m) {
  constexpr float s[9] = {0.f, 0.64278760968653933f, 0.98480775301220806f,
                          0.86602540378443865f, 0.34202014332566873f,
                          -0.34202014332566873f, -0.86602540378443865f,
                          -0.98480775301220806f, -0.64278760968653933f};
  return s[m];
}
template <>
__device__ __forceinline__ float wcos<8>(int m) {
  constexpr float c[8] = {1.f, 0.70710678118654752f, 0.f, -0.70710678118654752f,
                          -1.f, -0.70710678118654752f, 0.f, 0.70710678118654752f};
  return c[m];
}
template <>
__device__ __forceinline__ float wsin<8>(int m) {
  constexpr float s[8] = {0.f, 0.70710678118654752f, 1.f, 0.70710678118654752f,
                          0.f, -0.70710678118654752f, -1.f, -0.70710678118654752f};
  return s[m];
}

// Cooley-Tukey composite of size R1 * R2 in registers:
//   input n = R2 n1 + n2, output k = k1 + R1 k2, twiddle exp(-2 pi i n2 k1 / N)
template <int R1, int R2>
__device__ __forceinline__ void dft_composite(C2 (&v)[R1 * R2]) {
  constexpr int N = R1 * R2;
  C2 y[R2][R1];
#pragma unroll
  for (int n2 = 0; n2 < R2; ++n2) {
    C2 col[R1];
#pragma unroll
    for (int n1 = 0; n1 < R1; ++n1) col[n1] = v[R2 * n1 + n2];
    dft<R1>(col);
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) {
      const int m = (n2 * k1) % N;
      if (m == 0) {
        y[n2][k1] = col[k1];
      } else if (4 * m == N) {  // times -i: a rename plus one sign flip (ALU pipe)
        y[n2][k1] = {col[k1].im, col[k1].re ^ 0x8000000080000000ull};
      } else {
        y[n2][k1] = cmulc(col[k1], wcos<N>(m), -wsin<N>(m));
      }
    }
  }
#pragma unroll
  for (int k1 = 0; k1 < R1; ++k1) {
    C2 row[R2];
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) row[n2] = y[n2][k1];
    dft<R2>(row);
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = row[k2];
  }
}
template <>
__device__ __forceinline__ void dft<16>(C2 (&v)[16]) { dft_composite<4, 4>(v); }
template <>
__device__ __forceinline__ void dft<9>(C2 (&v)[9]) { dft_composite<3, 3>(v); }
template <>
__device__ __forceinline__ void dft<8>(C2 (&v)[8]) { dft_composite<4, 2>(v); }

// ---------------------------------------------------------------------------
constexpr int idem(int n2, int r) {  // CRT idempotent: 1 mod r, 0 mod n2 / r
  const int m = n2 / r;
  int x = m;
  while (x % r != 1 % r) x += m;
  return x % n2;
}

// TWG_: read the split twiddles from global memory (L1) instead of a per-CTA
// shared copy -- 11.5 KB less shared memory per CTA, which is what a fourth CTA
// per SM needs.
// DW_: two work buffers, alternating between items: stage A of item i + 1 may
// then start while slower warps are still in stage C of item i (one CTA
// barrier per item less, and a natural overlap of the two 5-warp stages).
template <int RA_, int RB_, int RC_, int G_, int NT_, int MINB_ = 2, int TWG_ = 0, int DW_ = 0>
struct Plan {
  static constexpr int RA = RA_, RB = RB_, RC = RC_, G = G_, NT = NT_, MINB = MINB_;
  static constexpr int TWG = TWG_, DW = DW_;
  static constexpr int N2 = RA * RB * RC, NK = N2 + 1, N = 2 * N2;
  static constexpr int SA = N2 / RA, SB = N2 / RB, SC = N2 / RC;  // Good's input map
  static constexpr int EA = idem(N2, RA), EB = idem(N2, RB), EC = idem(N2, RC);
  static constexpr int LB = RA * RC, LA = RC;  // in-place layout [RB][RA][RC]
  static constexpr int NTA = RB * RC, NTB = RA * RC;  // tasks per row pair
  // mirrored pairs {(kA, kB), (-kA, -kB)}: the self-paired ones are kA = 0
  // (RA odd) with kB = 0 and, for even RB, kB = RB / 2
  static constexpr int NSELF = (RA % 2 == 0 ? 2 : 1) * (RB % 2 == 0 ? 2 : 1);
  static constexpr int NTC = (RA * RB - NSELF) / 2 + NSELF;
  static constexpr int kRowBytes = N * 4;
  static constexpr int kStageBytes = 2 * G * kRowBytes;
  static constexpr int kWorkBytes = G * N2 * 16;
  static constexpr int kTwnBytes = TWG ? 0 : N2 * 16;
  static constexpr int kTaskBytes = ((NTC * 16 + 127) / 128) * 128;
  static constexpr int kWorkAll = (DW ? 2 : 1) * kWorkBytes;
  static constexpr int kSmem = kStageBytes + kWorkAll + kTwnBytes + kTaskBytes + 64;
  // thread slots per row pair, padded to half-warps so that the 64-bit shared
  // accesses of a half-warp stay inside one row pair (conflict-free strides)
  static constexpr int PTB = (NTB + 15) / 16 * 16, PTC = (NTC + 15) / 16 * 16;
  static_assert(G * PTC <= NT, "stage C keeps one task per thread (register accumulators)");
  static_assert(G * NTA <= NT && G * PTB <= NT, "one task per thread in every stage");
  static_assert(2 * G * NK * 4 <= kWorkBytes, "flush tile must fit the work buffer");
  static_assert(N2 % 2 == 0, "rows must be multiples of 16 bytes for TMA");
};

struct Params {
  const float* x;       // [nfield][nrow][N]
  float* out;           // modes 0/1: [nslot][nrow][NK];  mode 2: partial [njob][NK]
  const float4* twn;    // [N2] (wr, wr, wi, wi) of exp(-2 pi i p / N)
  const int4* ctask;    // [NTC] {offset of (kA,kB), offset of its mirror, k0, self}
  const float* scale;   // [2 G ngroup] per-row factor, 0 beyond nrow
  int64_t nslot;        // output slots (fields are slot-minor: field = ti * nslot + slot)
  int32_t ntimes;       // fields per slot
  int32_t nrow;
  int32_t ngroup;       // groups of G row pairs per field
  int32_t gpc;          // groups per chunk (1 in modes 0 / 1)
  int32_t nchunk;       // chunks per field
  int64_t njob;         // nslot * nchunk
  int32_t mode;         // 0 store, 1 add to out, 2 latitude-reduced partials
};

// position in the job / item sequence of one CTA (uniform across the CTA)
struct Cursor {
  int64_t job;
  int32_t ti, gi, gbeg, gend;
  int64_t slot;
  __device__ __forceinline__ void open(const Params& p) {
    if (job < p.njob) {
      slot = job / p.nchunk;
      const int chunk = static_cast<int>(job - slot * p.nchunk);
      gbeg = chunk * p.gpc;
      gend = min(p.ngroup, gbeg + p.gpc);
      gi = gbeg;
      ti = 0;
    }
  }
  __device__ __forceinline__ bool valid(const Params& p) const { return job < p.njob; }
  __device__ __forceinline__ bool last_of_job(const Params& p) const {
    return gi == gend - 1 && ti == p.ntimes - 1;
  }
  __device__ __forceinline__ void advance(const Params& p) {
    if (++gi == gend) {
      gi = gbeg;
      if (++ti == p.ntimes) {
        job += gridDim.x;
        open(p);
      }
    }
  }
};

template <class P>
__device__ __forceinline__ void issue_item(const Params& p, const Cursor& c, float* staging,
                                           uint64_t* bar) {
  const int64_t field = int64_t(c.ti) * p.nslot + c.slot;
  const float* base = p.x + field * p.nrow * int64_t(P::N);
  mbar_arrive_expect_tx(bar, P::kStageBytes);
#pragma unroll 1
  for (int r = 0; r < 2 * P::G; ++r) {
    const int row = min(2 * P::G * c.gi + r, p.nrow - 1);  // padding rows repeat the last row
    tma_load_1d(staging + r * P::N, base + int64_t(row) * P::N, P::kRowBytes, bar);
  }
}

template <class P, int MODE>
__global__ void __launch_bounds__(P::NT, P::MINB) spectrum_pfa_kernel(const Params p) {
  constexpr int RA = P::RA, RB = P::RB, RC = P::RC, G = P::G, NT = P::NT;
  constexpr int N2 = P::N2, NK = P::NK;
  extern __shared__ __align__(128) unsigned char smem[];
  float* staging = reinterpret_cast<float*>(smem);
  // in-place work array, planar: re and im as separate 64-bit planes (an
  // interleaved 128-bit element would need its four registers adjacent, which
  // costs four MOVs per store)
  u64* work_re = reinterpret_cast<u64*>(smem + P::kStageBytes);
  u64* work_im = work_re + G * N2;
  u64* const work_base = work_re;
  const ulonglong2* twn =
      P::TWG ? reinterpret_cast<const ulonglong2*>(p.twn)
             : reinterpret_cast<const ulonglong2*>(smem + P::kStageBytes + P::kWorkAll);
  const int4* ctask = reinterpret_cast<const int4*>(smem + P::kStageBytes + P::kWorkAll +
                                                     P::kTwnBytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + P::kStageBytes + P::kWorkAll +
                                               P::kTwnBytes + P::kTaskBytes);
  float* tile = reinterpret_cast<float*>(work_re);  // flush tile [2 G][NK], aliases `work`
  uint32_t wsel = 0;  // DW: which work buffer the current item uses
  const int tid = threadIdx.x;

  {
    if (!P::TWG) {
      float4* d = reinterpret_cast<float4*>(smem + P::kStageBytes + P::kWorkAll);
      for (int i = tid; i < N2; i += NT) d[i] = p.twn[i];
    }
    int4* t = reinterpret_cast<int4*>(smem + P::kStageBytes + P::kWorkAll + P::kTwnBytes);
    for (int i = tid; i < P::NTC; i += NT) t[i] = p.ctask[i];
  }
  Cursor cur;
  cur.job = blockIdx.x;
  cur.open(p);
  if (tid == 0) {
    mbar_init(full, 1);
    mbar_fence_init();
    if (cur.valid(p)) issue_item<P>(p, cur, staging, full);
  }
  __syncthreads();

  // ---- thread-invariant task geometry --------------------------------------
  // stage A (radix RA, staging -> work): half-warps share nC and span nB, so the
  // gathered LDS.64 hit 16 distinct bank pairs (SB is odd)
  const bool a_on = tid < G * P::NTA;
  int a_e0 = 0, a_src = 0, a_dst = 0;
  {
    const int g = tid / P::NTA, r = tid - g * P::NTA;
    const int nC = r / RB, nB = r - nC * RB;
    a_e0 = (P::SB * nB + P::SC * nC) % N2;
    a_src = 2 * g * N2;  // float2 index of row A of the pair; row B follows at + N2
    a_dst = g * N2 + nB * P::LB + nC;
  }
  // stage B (radix RB, in place): consecutive threads, consecutive elements
  bool b_on;
  int b_off = 0;
  {
    const int g = tid / P::PTB, r = tid - g * P::PTB;
    b_on = g < G && r < P::NTB;
    b_off = g * N2 + r;
  }
  // stage C (radix RC on a butterfly and its mirror, then the real-input split)
  int c_g = tid / P::PTC, c_off1 = 0, c_off2 = 0, c_k0 = 0, c_self = 0;
  const bool c_on = c_g < G && tid - c_g * P::PTC < P::NTC;
  if (c_on) {
    const int4 t = ctask[tid - c_g * P::PTC];
    c_off1 = c_g * N2 + t.x;
    c_off2 = c_g * N2 + t.y;
    c_k0 = t.z;
    c_self = t.w;
  }
  u64 acc_a[RC], acc_b[RC];
#pragma unroll
  for (int i = 0; i < RC; ++i) acc_a[i] = acc_b[i] = 0ull;

  uint32_t parity = 0;
  while (cur.valid(p)) {
    Cursor nxt = cur;
    nxt.advance(p);
    if (P::DW) {
      work_re = work_base + wsel * (P::kWorkBytes / 8);
      work_im = work_re + G * N2;
      tile = reinterpret_cast<float*>(work_re);
      wsel ^= 1u;
    }
    mbar_wait(full, parity);
    parity ^= 1u;

    // ---- stage A --------------------------------------------------------------
    if (a_on) {
      const float2* ra = reinterpret_cast<const float2*>(staging) + a_src;
      const float2* rb = ra + N2;
      C2 v[RA];
#pragma unroll
      for (int nA = 0; nA < RA; ++nA) {
        int e = a_e0 + P::SA * nA;
        e = e >= N2 ? e - N2 : e;
        const float2 a = ra[e], b = rb[e];
        v[nA].re = pk2v(a.x, b.x);
        v[nA].im = pk2v(a.y, b.y);
      }
      dft<RA>(v);
#pragma unroll
      for (int kA = 0; kA < RA; ++kA) {
        work_re[a_dst + kA * P::LA] = v[kA].re;
        work_im[a_dst + kA * P::LA] = v[kA].im;
      }
    }
    __syncthreads();  // staging consumed, work complete
    if (tid == 0 && nxt.valid(p)) issue_item<P>(p, nxt, staging, full);

    // ---- stage B --------------------------------------------------------------
    if (b_on) {
      C2 v[RB];
#pragma unroll
      for (int nB = 0; nB < RB; ++nB)
        v[nB] = {work_re[b_off + nB * P::LB], work_im[b_off + nB * P::LB]};
      dft<RB>(v);
#pragma unroll
      for (int kB = 0; kB < RB; ++kB) {
        work_re[b_off + kB * P::LB] = v[kB].re;
        work_im[b_off + kB * P::LB] = v[kB].im;
      }
    }
    __syncthreads();

    // ---- stage C + split + power ---------------------------------------------
    if (c_on) {
      C2 b1[RC], b2[RC];
#pragma unroll
      for (int i = 0; i < RC; ++i) {
        b1[i] = {work_re[c_off1 + i], work_im[c_off1 + i]};
        b2[i] = {work_re[c_off2 + i], work_im[c_off2 + i]};
      }
      dft<RC>(b1);
      dft<RC>(b2);
      u64 scl = 0ull;
      if (MODE == 2) {
        const int row = 2 * (G * cur.gi + c_g);
        scl = pk2(__ldg(p.scale + row), __ldg(p.scale + row + 1));
      }
#pragma unroll
      for (int kC = 0; kC < RC; ++kC) {
        int pbin = c_k0 + (kC * P::EC) % N2;
        pbin = pbin >= N2 ? pbin - N2 : pbin;
        const ulonglong2 w = twn[pbin];  // (wr, wr), (wi, wi)
        const C2 zp = b1[kC], zq = b2[(RC - kC) % RC];
        const u64 e_re = add2(zp.re, zq.re), e_im = sub2(zp.im, zq.im);
        const u64 d_re = sub2(zp.re, zq.re), d_im = add2(zp.im, zq.im);
        // W (-i D):  re = wr d_im + wi d_re,  im = wi d_im - wr d_re
        const u64 t1 = fma2(w.y, d_re, mul2(w.x, d_im));
        const u64 t2 = sub2(mul2(w.x, d_re), mul2(w.y, d_im));  // = -im
        const u64 xa_re = add2(e_re, t1), xa_im = sub2(e_im, t2);
        const u64 xb_re = sub2(e_re, t1), xb_im = add2(e_im, t2);
        if (MODE == 2) {
          const u64 pa = fma2(xa_im, xa_im, mul2(xa_re, xa_re));
          const u64 pb = fma2(xb_im, xb_im, mul2(xb_re, xb_re));
          acc_a[kC] = fma2(pa, scl, acc_a[kC]);
          acc_b[kC] = fma2(pb, scl, acc_b[kC]);
        } else {
          acc_a[kC] = fma2(xa_im, xa_im, fma2(xa_re, xa_re, acc_a[kC]));
          acc_b[kC] = fma2(xb_im, xb_im, fma2(xb_re, xb_re, acc_b[kC]));
        }
      }
    }

    if (cur.last_of_job(p)) {
      __syncthreads();  // everyone is done reading `work`; it becomes the tile
      if (c_on) {
        float* ta = tile + (2 * c_g) * NK;
        float* tb = ta + NK;
        // modes 0 / 1: the per-row factor is applied here, so that the copy-out
        // below is a flat, unrolled add (mode 2 applied it per item already)
        float sa = 1.f, sb = 1.f;
        if (MODE != 2) {
          const int row = 2 * (G * cur.gi + c_g);
          sa = __ldg(p.scale + row);
          sb = __ldg(p.scale + row + 1);
        }
#pragma unroll
        for (int kC = 0; kC < RC; ++kC) {
          int pbin = c_k0 + (kC * P::EC) % N2;
          pbin = pbin >= N2 ? pbin - N2 : pbin;
          const bool va = !c_self || kC <= (RC - kC) % RC;
          const bool vb = va && (2 * pbin != N2);
          const float h0 = pbin == 0 ? 0.5f : 1.f;  // c_k = 1 for the zonal mean
          if (va) {
            ta[pbin] = lo2(acc_a[kC]) * (sa * h0);
            tb[pbin] = hi2(acc_a[kC]) * (sb * h0);
          }
          if (vb) {
            ta[N2 - pbin] = lo2(acc_b[kC]) * sa;
            tb[N2 - pbin] = hi2(acc_b[kC]) * sb;
          }
          acc_a[kC] = acc_b[kC] = 0ull;
        }
      }
      __syncthreads();
      if (MODE == 2) {
        // rows already carry their weights: add the 2 G rows in a fixed order
        float* o = p.out + cur.job * NK;
        for (int k = tid; k < NK; k += NT) {
          float s = 0.f;
#pragma unroll
          for (int r = 0; r < 2 * G; ++r) s += tile[r * NK + k];
          o[k] = s;
        }
      } else {
        // the rows of a group are contiguous in `out`: one flat range; loads of
        // the read-modify-write are issued kU at a time (the un-unrolled loop
        // held a third of the kernel's stall samples on the first FFMA)
        const int row0 = 2 * G * cur.gi;
        const int n = min(2 * G, p.nrow - row0) * NK;
        float* o = p.out + (cur.slot * p.nrow + row0) * int64_t(NK);
        constexpr int kU = 6;
        for (int i0 = tid; i0 < n; i0 += kU * NT) {
          float v[kU];
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            const int i = i0 + u * NT;
            v[u] = (MODE == 1 && i < n) ? o[i] : 0.f;
          }
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            const int i = i0 + u * NT;
            if (i < n) o[i] = v[u] + tile[i];
          }
        }
      }
    }
    // one work buffer: it (and the tile in it) is overwritten by the next stage
    // A; two: the next item writes the other one, and the one after that is
    // behind the two barriers of the next item
    if (!P::DW) __syncthreads();
    cur = nxt;
  }
}

// ---------------------------------------------------------------------------
// Warp-specialised variant.  The kernel above runs the three stages one after
// the other behind CTA-wide barriers; ncu (profiles/r2_k4_*): a third of its
// stall samples sit on those barriers, because stage B has work for only 3 of
// the 5 warps and the TMA wait idles everyone.  Here the stages are THREE WARP
// GROUPS of one persistent CTA per SM that work on different items at the same
// time and hand buffers over through mbarriers:
//     TMA -> staging[2] -> group A (radix RA) -> work[3] -> group B (radix RB,
//     in place) -> group C (radix RC on mirrored pairs, split, power, accumulate)
// Group sizes follow the per-item instruction counts (5 : 3 : 5 warps for
// 720 = 9 * 16 * 5), so every group is busy all the time and the pipes see A's
// shared-memory traffic, B's FADD2s and C's mixed work interleaved.  Within a
// group: a named barrier (bar.sync id, nthreads); between groups: mbarrier
// arrive (release) by one elected thread after the group's barrier, try_wait
// (acquire) by the consumers.
// ---------------------------------------------------------------------------
template <class P>
struct WsLayout {
  static constexpr int NSTG = 2, NWRK = 3;
  static constexpr int WA = (P::G * P::NTA + 31) / 32, WB = (P::G * P::PTB + 31) / 32,
                       WC = (P::G * P::PTC + 31) / 32;
  static constexpr int NT = 32 * (WA + WB + WC);
  static constexpr int kTileBytes = ((2 * P::G * P::NK * 4 + 127) / 128) * 128;
  static constexpr int kOffWork = NSTG * P::kStageBytes;
  static constexpr int kOffTwn = kOffWork + NWRK * P::kWorkBytes;
  static constexpr int kOffTask = kOffTwn + P::kTwnBytes;
  static constexpr int kOffTile = kOffTask + P::kTaskBytes;
  static constexpr int kOffBar = kOffTile + kTileBytes;
  static constexpr int kSmem = kOffBar + 128;
};

__device__ __forceinline__ void group_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <class P, int MODE>
__global__ void __launch_bounds__(WsLayout<P>::NT, 1) spectrum_pfa_ws_kernel(const Params p) {
  using L = WsLayout<P>;
  constexpr int RA = P::RA, RB = P::RB, RC = P::RC, G = P::G;
  constexpr int N2 = P::N2, NK = P::NK;
  constexpr int NTA_T = 32 * L::WA, NTB_T = 32 * L::WB, NTC_T = 32 * L::WC;
  extern __shared__ __align__(128) unsigned char smem[];
  const ulonglong2* twn = reinterpret_cast<const ulonglong2*>(smem + L::kOffTwn);
  const int4* ctask = reinterpret_cast<const int4*>(smem + L::kOffTask);
  float* tile = reinterpret_cast<float*>(smem + L::kOffTile);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* stg_full = bars;                  // [NSTG]  TMA landed
  uint64_t* ab_full = bars + L::NSTG;         // [NWRK]  stage A wrote work[w]
  uint64_t* bc_full = ab_full + L::NWRK;      // [NWRK]  stage B finished work[w]
  uint64_t* w_empty = bc_full + L::NWRK;      // [NWRK]  stage C has read work[w]
  const int tid = threadIdx.x;
  const int warp = tid >> 5;

  {
    float4* d = reinterpret_cast<float4*>(smem + L::kOffTwn);
    for (int i = tid; i < N2; i += L::NT) d[i] = p.twn[i];
    int4* t = reinterpret_cast<int4*>(smem + L::kOffTask);
    for (int i = tid; i < P::NTC; i += L::NT) t[i] = p.ctask[i];
  }
  if (tid == 0) {
    for (int i = 0; i < L::NSTG + 3 * L::NWRK; ++i) mbar_init(bars + i, 1);
    mbar_fence_init();
  }
  __syncthreads();

  Cursor cur;
  cur.job = blockIdx.x;
  cur.open(p);

  if (warp < L::WA) {
    // ===================== group A: staging -> work (radix RA) =================
    const int t = tid;
    const bool on = t < G * P::NTA;
    const int g = t / P::NTA, r = t - g * P::NTA;
    const int nC = r / RB, nB = r - nC * RB;
    const int e0 = (P::SB * nB + P::SC * nC) % N2;
    const int src0 = 2 * g * N2, dst0 = g * N2 + nB * P::LB + nC;
    Cursor pre = cur;  // the item whose rows are fetched next
    if (t == 0) {
      for (int k = 0; k < L::NSTG && pre.valid(p); ++k) {
        issue_item<P>(p, pre, reinterpret_cast<float*>(smem + k * P::kStageBytes), &stg_full[k]);
        pre.advance(p);
      }
    } else {
      for (int k = 0; k < L::NSTG && pre.valid(p); ++k) pre.advance(p);
    }
    for (uint32_t it = 0; cur.valid(p); ++it, cur.advance(p)) {
      const int s = it % L::NSTG, w = it % L::NWRK;
      mbar_wait(&stg_full[s], (it / L::NSTG) & 1u);
      mbar_wait(&w_empty[w], ((it / L::NWRK) & 1u) ^ 1u);
      if (on) {
        const float2* ra = reinterpret_cast<const float2*>(smem + s * P::kStageBytes) + src0;
        const float2* rb = ra + N2;
        u64* wre = reinterpret_cast<u64*>(smem + L::kOffWork + w * P::kWorkBytes);
        u64* wim = wre + G * N2;
        C2 v[RA];
#pragma unroll
        for (int nA = 0; nA < RA; ++nA) {
          int e = e0 + P::SA * nA;
          e = e >= N2 ? e - N2 : e;
          const float2 a = ra[e], b = rb[e];
          v[nA].re = pk2v(a.x, b.x);
          v[nA].im = pk2v(a.y, b.y);
        }
        dft<RA>(v);
#pragma unroll
        for (int kA = 0; kA < RA; ++kA) {
          wre[dst0 + kA * P::LA] = v[kA].re;
          wim[dst0 + kA * P::LA] = v[kA].im;
        }
      }
      group_sync(1, NTA_T);
      if (t == 0) {
        mbar_arrive(&ab_full[w]);
        if (pre.valid(p))  // staging[s] is free again: fetch item it + NSTG into it
          issue_item<P>(p, pre, reinterpret_cast<float*>(smem + s * P::kStageBytes), &stg_full[s]);
      }
      if (pre.valid(p)) pre.advance(p);
    }
  } else if (warp < L::WA + L::WB) {
    // ===================== group B: radix RB, in place =========================
    const int t = tid - NTA_T;
    const int g = t / P::PTB, r = t - g * P::PTB;
    const bool on = g < G && r < P::NTB;
    const int off = g * N2 + r;
    for (uint32_t it = 0; cur.valid(p); ++it, cur.advance(p)) {
      const int w = it % L::NWRK;
      mbar_wait(&ab_full[w], (it / L::NWRK) & 1u);
      if (on) {
        u64* wre = reinterpret_cast<u64*>(smem + L::kOffWork + w * P::kWorkBytes);
        u64* wim = wre + G * N2;
        C2 v[RB];
#pragma unroll
        for (int nB = 0; nB < RB; ++nB) v[nB] = {wre[off + nB * P::LB], wim[off + nB * P::LB]};
        dft<RB>(v);
#pragma unroll
        for (int kB = 0; kB < RB; ++kB) {
          wre[off + kB * P::LB] = v[kB].re;
          wim[off + kB * P::LB] = v[kB].im;
        }
      }
      group_sync(2, NTB_T);
      if (t == 0) mbar_arrive(&bc_full[w]);
    }
  } else {
    // ========== group C: radix RC on mirrored pairs, split, power, flush =======
    const int t = tid - NTA_T - NTB_T;
    const int c_g = t / P::PTC;
    const bool on = c_g < G && t - c_g * P::PTC < P::NTC;
    int c_off1 = 0, c_off2 = 0, c_k0 = 0, c_self = 0;
    if (on) {
      const int4 q = ctask[t - c_g * P::PTC];
      c_off1 = c_g * N2 + q.x;
      c_off2 = c_g * N2 + q.y;
      c_k0 = q.z;
      c_self = q.w;
    }
    u64 acc_a[RC], acc_b[RC];
#pragma unroll
    for (int i = 0; i < RC; ++i) acc_a[i] = acc_b[i] = 0ull;
    for (uint32_t it = 0; cur.valid(p); ++it, cur.advance(p)) {
      const int w = it % L::NWRK;
      mbar_wait(&bc_full[w], (it / L::NWRK) & 1u);
      C2 b1[RC], b2[RC];
      if (on) {
        const u64* wre = reinterpret_cast<const u64*>(smem + L::kOffWork + w * P::kWorkBytes);
        const u64* wim = wre + G * N2;
#pragma unroll
        for (int i = 0; i < RC; ++i) {
          b1[i] = {wre[c_off1 + i], wim[c_off1 + i]};
          b2[i] = {wre[c_off2 + i], wim[c_off2 + i]};
        }
      }
      group_sync(3, NTC_T);  // everyone has its values: the buffer can go back to A
      if (t == 0) mbar_arrive(&w_empty[w]);
      if (on) {
        dft<RC>(b1);
        dft<RC>(b2);
        u64 scl = 0ull;
        if (MODE == 2) {
          const int row = 2 * (G * cur.gi + c_g);
          scl = pk2(__ldg(p.scale + row), __ldg(p.scale + row + 1));
        }
#pragma unroll
        for (int kC = 0; kC < RC; ++kC) {
          int pbin = c_k0 + (kC * P::EC) % N2;
          pbin = pbin >= N2 ? pbin - N2 : pbin;
          const ulonglong2 wv = twn[pbin];  // (wr, wr), (wi, wi)
          const C2 zp = b1[kC], zq = b2[(RC - kC) % RC];
          const u64 e_re = add2(zp.re, zq.re), e_im = sub2(zp.im, zq.im);
          const u64 d_re = sub2(zp.re, zq.re), d_im = add2(zp.im, zq.im);
          const u64 t1 = fma2(wv.y, d_re, mul2(wv.x, d_im));
          const u64 t2 = sub2(mul2(wv.x, d_re), mul2(wv.y, d_im));
          const u64 xa_re = add2(e_re, t1), xa_im = sub2(e_im, t2);
          const u64 xb_re = sub2(e_re, t1), xb_im = add2(e_im, t2);
          if (MODE == 2) {
            const u64 pa = fma2(xa_im, xa_im, mul2(xa_re, xa_re));
            const u64 pb = fma2(xb_im, xb_im, mul2(xb_re, xb_re));
            acc_a[kC] = fma2(pa, scl, acc_a[kC]);
            acc_b[kC] = fma2(pb, scl, acc_b[kC]);
          } else {
            acc_a[kC] = fma2(xa_im, xa_im, fma2(xa_re, xa_re, acc_a[kC]));
            acc_b[kC] = fma2(xb_im, xb_im, fma2(xb_re, xb_re, acc_b[kC]));
          }
        }
      }
      if (cur.last_of_job(p)) {
        // the flush tile is private to this group
        if (on) {
          float* ta = tile + (2 * c_g) * NK;
          float* tb = ta + NK;
          float sa = 1.f, sb = 1.f;
          if (MODE != 2) {
            const int row = 2 * (G * cur.gi + c_g);
            sa = __ldg(p.scale + row);
            sb = __ldg(p.scale + row + 1);
          }
#pragma unroll
          for (int kC = 0; kC < RC; ++kC) {
            int pbin = c_k0 + (kC * P::EC) % N2;
            pbin = pbin >= N2 ? pbin - N2 : pbin;
            const bool va = !c_self || kC <= (RC - kC) % RC;
            const bool vb = va && (2 * pbin != N2);
            const float h0 = pbin == 0 ? 0.5f : 1.f;
            if (va) {
              ta[pbin] = lo2(acc_a[kC]) * (sa * h0);
              tb[pbin] = hi2(acc_a[kC]) * (sb * h0);
            }
            if (vb) {
              ta[N2 - pbin] = lo2(acc_b[kC]) * sa;
              tb[N2 - pbin] = hi2(acc_b[kC]) * sb;
            }
            acc_a[kC] = acc_b[kC] = 0ull;
          }
        }
        group_sync(3, NTC_T);
        if (MODE == 2) {
          float* o = p.out + cur.job * NK;
          for (int k = t; k < NK; k += NTC_T) {
            float sum = 0.f;
#pragma unroll
            for (int rr = 0; rr < 2 * G; ++rr) sum += tile[rr * NK + k];
            o[k] = sum;
          }
        } else {
          const int row0 = 2 * G * cur.gi;
          const int n = min(2 * G, p.nrow - row0) * NK;
          float* o = p.out + (cur.slot * p.nrow + row0) * int64_t(NK);
          constexpr int kU = 6;
          for (int i0 = t; i0 < n; i0 += kU * NTC_T) {
            float v[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
              const int i = i0 + u * NTC_T;
              v[u] = (MODE == 1 && i < n) ? o[i] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
              const int i = i0 + u * NTC_T;
              if (i < n) o[i] = v[u] + tile[i];
            }
          }
        }
        group_sync(3, NTC_T);  // the tile is reused by the next job
      }
    }
  }
}

// sums the per-chunk partials of mode 2 in a fixed order
__global__ void latsum_finalize_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                       int64_t nslot, int nchunk, int nk, int accumulate) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= nslot * nk) return;
  const int64_t slot = i / nk;
  const int k = static_cast<int>(i - slot * nk);
  float s = 0.f;
  for (int c = 0; c < nchunk; ++c) s += partial[(slot * nchunk + c) * nk + k];
  out[i] = accumulate ? out[i] + s : s;
}

template <class P>
static void build_tables(std::vector<float4>* twn, std::vector<int4>* ctask) {
  const double two_pi = 6.283185307179586476925286766559;
  twn->resize(P::N2);
  for (int k = 0; k < P::N2; ++k) {
    const double a = -two_pi * k / P::N;
    const float wr = static_cast<float>(cos(a)), wi = static_cast<float>(sin(a));
    (*twn)[k] = make_float4(wr, wr, wi, wi);
  }
  ctask->clear();
  std::vector<char> seen(P::RA * P::RB, 0);
  // kA = 1, 2, ... first: their kB runs are whole half-warps (16 consecutive kB
  // hit 16 distinct bank pairs because LB is odd); the shorter kA = 0 run last
  for (int ia = 1; ia <= P::RA; ++ia)
    for (int kB = 0; kB < P::RB; ++kB) {
      const int kA = ia % P::RA;
      if (seen[kA * P::RB + kB]) continue;
      const int qA = (P::RA - kA) % P::RA, qB = (P::RB - kB) % P::RB;
      seen[kA * P::RB + kB] = 1;
      seen[qA * P::RB + qB] = 1;
      int4 t;
      t.x = kB * P::LB + kA * P::LA;
      t.y = qB * P::LB + qA * P::LA;
      t.z = (kA * P::EA + kB * P::EB) % P::N2;
      t.w = (qA == kA && qB == kB) ? 1 : 0;
      ctask->push_back(t);
    }
}

template <class P, bool warp_specialised = false>
static int launch(wb2_ctx* ctx, const float* x, int64_t nfield, int32_t nrow,
                  const double* scale, float* out, int mode, int64_t nslot) {
  const int latsum_accumulate = mode == 3;  // mode 3 = mode 2 adding to `out`
  if (mode == 3) mode = 2;
  std::vector<float4> twn;
  std::vector<int4> ctask;
  build_tables<P>(&twn, &ctask);
  if (static_cast<int>(ctask.size()) != P::NTC) {
    set_error("spectrum_pfa: task table has %zu entries, expected %d", ctask.size(), P::NTC);
    return WB2_EINVAL;
  }
  Params p;
  p.x = x;
  p.nslot = nslot;
  p.ntimes = static_cast<int32_t>(nfield / nslot);
  p.nrow = nrow;
  const int npair = (nrow + 1) / 2;
  p.ngroup = (npair + P::G - 1) / P::G;
  p.mode = mode;
  // latitude chunks of mode 2: enough jobs to fill the machine, few partials
  const int64_t want_jobs = int64_t(ctx->num_sms) * (warp_specialised ? 1 : P::MINB) * 4;
  if (mode == 2) {
    int nchunk = static_cast<int>((want_jobs + nslot - 1) / nslot);
    if (nchunk < 1) nchunk = 1;
    if (nchunk > p.ngroup) nchunk = p.ngroup;
    p.gpc = (p.ngroup + nchunk - 1) / nchunk;
    p.nchunk = (p.ngroup + p.gpc - 1) / p.gpc;
  } else {
    p.gpc = 1;
    p.nchunk = p.ngroup;
  }
  p.njob = nslot * p.nchunk;
  // per-row factor: circumference / N^2 (rfft norm='forward') * 2 (c_k; bin 0 is
  // halved at the flush) / 4 (the split computes 2 X_p): scale / (2 N^2)
  std::vector<float> sc(size_t(2) * P::G * p.ngroup, 0.f);
  for (int i = 0; i < nrow; ++i)
    sc[i] = static_cast<float>(scale[i] / (2.0 * double(P::N) * double(P::N)));

  Packer pk(ctx);
  const size_t o1 = pk.add(twn.data(), twn.size() * sizeof(float4));
  const size_t o2 = pk.add(ctask.data(), ctask.size() * sizeof(int4));
  const size_t o3 = pk.add(sc.data(), sc.size() * sizeof(float));
  size_t o4 = 0;
  if (mode == 2) o4 = pk.reserve(size_t(p.njob) * P::NK * sizeof(float));
  WB2_TRY(pk.commit());
  p.twn = pk.dev<float4>(o1);
  p.ctask = pk.dev<int4>(o2);
  p.scale = pk.dev<float>(o3);
  p.out = mode == 2 ? pk.dev<float>(o4) : out;

  if constexpr (warp_specialised) {
    using L = WsLayout<P>;
    const int64_t max_cta = ctx->num_sms;  // one persistent CTA per SM
    const unsigned grid = static_cast<unsigned>(p.njob < max_cta ? p.njob : max_cta);
    auto kernel = mode == 0 ? spectrum_pfa_ws_kernel<P, 0>
                            : (mode == 1 ? spectrum_pfa_ws_kernel<P, 1>
                                         : spectrum_pfa_ws_kernel<P, 2>);
    WB2_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      L::kSmem));
    kernel<<<grid, L::NT, L::kSmem, ctx->stream>>>(p);
  } else {
    const int64_t max_cta = int64_t(ctx->num_sms) * P::MINB;
    const unsigned grid = static_cast<unsigned>(p.njob < max_cta ? p.njob : max_cta);
    auto kernel = mode == 0 ? spectrum_pfa_kernel<P, 0>
                            : (mode == 1 ? spectrum_pfa_kernel<P, 1> : spectrum_pfa_kernel<P, 2>);
    WB2_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      P::kSmem));
    kernel<<<grid, P::NT, P::kSmem, ctx->stream>>>(p);
  }
  WB2_CUDA_TRY(cudaGetLastError());
  ctx->launches += 1;
  if (mode == 2) {
    const int64_t n = nslot * P::NK;
    latsum_finalize_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(
        p.out, out, nslot, p.nchunk, P::NK, latsum_accumulate);
    WB2_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
  }
  WB2_TRY(pk.release());
  return WB2_OK;
}

}  // namespace pfa

// Returns 1 when the prime-factor kernel handled the call, 0 when the shape /
// alignment is not eligible (the caller falls back), < 0 on error.
//   mode 0: out[field][row][k] = S;   mode 1: out[slot][row][k] += sum_time S;
//   mode 2: out[slot][k] = sum_time sum_row scale[row] * S-without-circumference
//           (`scale` already holds circumference * row weight);  mode 3: same,
//           added to `out`.
int spectrum_pfa_try(wb2_ctx* ctx, const float* x, int64_t nfield, int32_t nrow, int32_t ncol,
                     const double* scale, float* out, int mode, int64_t nslot) {
  const char* force = getenv("WB2_SPECTRUM_PATH");
  if (force && strcmp(force, "pfa") != 0) return 0;
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return 0;
  if (nrow < 1 || nfield < 1 || nslot < 1 || nfield % nslot != 0) return 0;
  int rc;
  switch (ncol) {
    case 1440: {
      const char* plan = getenv("WB2_PFA_PLAN");  // experiments: CTA shape
      if (plan && plan[0] == 'w')  // warp-specialised pipeline, one CTA per SM
        rc = pfa::launch<pfa::Plan<9, 16, 5, 2, 160, 3>, true>(ctx, x, nfield, nrow, scale, out,
                                                               mode, nslot);
      else if (plan && plan[0] == '4')  // 4 CTAs per SM (102 registers, twiddles via L1)
        rc = pfa::launch<pfa::Plan<9, 16, 5, 2, 160, 4, 1>>(ctx, x, nfield, nrow, scale, out,
                                                            mode, nslot);
      else if (plan && plan[0] == '0')
        rc = pfa::launch<pfa::Plan<9, 16, 5, 3, 256, 2>>(ctx, x, nfield, nrow, scale, out, mode,
                                                         nslot);
      else if (plan && plan[0] == '2')
        rc = pfa::launch<pfa::Plan<9, 16, 5, 1, 96, 5>>(ctx, x, nfield, nrow, scale, out, mode,
                                                        nslot);
      else if ((plan && plan[0] == '1') || mode == 0)
        // per-time output: 3 CTAs of 5 warps per SM (measured best)
        rc = pfa::launch<pfa::Plan<9, 16, 5, 2, 160, 3>>(ctx, x, nfield, nrow, scale, out, mode,
                                                         nslot);
      else if (plan && plan[0] == 'd')  // double work buffer, 3 CTAs, twiddles via L1
        rc = pfa::launch<pfa::Plan<9, 16, 5, 2, 160, 3, 1, 1>>(ctx, x, nfield, nrow, scale, out,
                                                               mode, nslot);
      else  // time sum / latitude reduction: 4 CTAs per SM measured 4-7 % faster
        rc = pfa::launch<pfa::Plan<9, 16, 5, 2, 160, 4, 1>>(ctx, x, nfield, nrow, scale, out,
                                                            mode, nslot);
      break;
    }
    case 720: rc = pfa::launch<pfa::Plan<9, 8, 5, 5, 256>>(ctx, x, nfield, nrow, scale, out,
                                                            mode, nslot); break;
    case 240: rc = pfa::launch<pfa::Plan<3, 8, 5, 6, 256>>(ctx, x, nfield, nrow, scale, out,
                                                            mode, nslot); break;
    default: return 0;
  }
  return rc == WB2_OK ? 1 : rc;
}

}  // namespace wb2
