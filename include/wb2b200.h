/* wb2b200.h -- C ABI of the B200-native WeatherBench2 hot path (libwb2b200.so).
 *
 * The reference (google-research/weatherbench2) is pure Python and has no FFI of
 * its own; these entry points are what its operator classes bind through
 * ctypes.  Each entry point cites the reference interface
 * it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every function returns 0 on success, a negative WB2_E* code on failure and
 *     never throws/aborts; wb2_last_error() gives the message (thread-local);
 *   - the caller owns every data buffer.  "dev" pointers are CUDA device
 *     pointers on the context's device, "host" pointers are ordinary host
 *     memory (pinned or pageable);  small descriptor arrays (weights, offset
 *     tables) are ALWAYS host pointers and are copied by the library;
 *   - kernels are enqueued on the context's stream; results written to device
 *     memory are ordered on that stream (wb2_synchronize / wb2_memcpy_d2h to
 *     read them);  the *_host entry points are synchronous end to end;
 *   - outputs are raw weighted SUMS (and weight sums) in float64; division,
 *     sqrt, ACC ratio, CRPS = skill - spread/2 stay in the Python operator so
 *     the NaN / zero-weight rules remain visible (weatherbench2/metrics.py:161);
 *   - reductions are deterministic (fixed-order tree, no float atomics).
 *
 * Field addressing.  A "field" is one 2-D (row, col) slab with `col`
 * contiguous: (latitude, longitude) for raw 0.25-degree data or
 * (longitude, latitude) for the reference's mock / regridded layout
 * (weatherbench2/schema.py:83).  Operands are addressed by a base pointer plus
 * a per-field element-offset table, so broadcasting (truth without lead_time),
 * the by-init gather `truth.sel(time=valid_time)` (evaluation.py:475) and the
 * climatology day-of-year lookup (metrics.py:398-404) need no materialised
 * copies.
 */
#ifndef WB2B200_H_
#define WB2B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WB2_VERSION 100 /* 0.1.0 */

enum {
  WB2_OK = 0,
  WB2_EINVAL = -1,   /* bad argument */
  WB2_ECUDA = -2,    /* CUDA runtime error (message has the cudaError string) */
  WB2_ENOMEM = -3,
  WB2_EUNSUPPORTED = -4
};

enum { WB2_F32 = 0, WB2_F64 = 1 };

#define WB2_MAX_REGIONS 32

/* number of per-(field, region) float64 outputs of wb2_det_metrics */
#define WB2_DET_NSTAT 10
/* [0] sum W*d^2   [1] sum W*|d|   [2] sum W*d          (d  = f - t)
 * [3] sum W*fa*ta [4] sum W*fa^2  [5] sum W*ta^2       (fa = f - c, ta = t - c)
 * [6] sum W*[d valid] [7] sum W*[fa*ta valid] [8] sum W*[fa valid]
 * [9] sum W*[ta valid]   -- the xarray `sum_of_weights` of each average     */

/* number of per-(field, region) float64 outputs of wb2_ens_metrics */
#define WB2_ENS_NSTAT 10
/* [0] sum W*skill_pt      skill_pt  = mean_m |t - x_m|        (metrics.py:824)
 * [1] sum W*spread_pt     spread_pt = 2 mean_m((2r_m-M-1)x_m)/(M-1)  (:805-813)
 * [2] sum W*(t - xbar)^2                                       (metrics.py:1330)
 * [3] sum W*var_ddof1                                          (metrics.py:1238)
 * [4] sum W*((t - xbar)^2 - var/M)                             (metrics.py:564-565)
 * [5..9] the matching sums of W*[value valid]                                */

typedef struct wb2_ctx wb2_ctx;

/* Region / latitude weights of one launch.  The weight of cell (row, col) for
 * region r is   W_r = row_w[r][row] * seg_w[r][seg(col)] * col_w[col] * cell_w[row][col]
 * where seg(col) = k for seg_start[k] <= col < seg_start[k+1].
 * This factorisation holds every reference region exactly
 * (weatherbench2/regions.py:57-158): latitude weights and latitude boxes go in
 * row_w (or col_w*seg_w for the latitude-fastest layout), longitude boxes in
 * seg_w (or row_w), a LandRegion mask in cell_w.
 * zero_skip != 0 reproduces `dataset.where(weights > 0, 0)` (metrics.py:160):
 * cells whose weight is zero contribute nothing, even when they hold NaN/Inf. */
typedef struct {
  int32_t nrow;            /* slow spatial axis length                         */
  int32_t ncol;            /* contiguous spatial axis length                   */
  int64_t row_stride;      /* elements between consecutive rows (>= ncol)      */
  int32_t nregion;         /* 1..WB2_MAX_REGIONS                               */
  int32_t nseg;            /* >= 1                                             */
  const double* row_w;     /* host [nregion][nrow]                             */
  const int32_t* seg_start;/* host [nseg + 1], seg_start[0]=0, [nseg]=ncol     */
  const double* seg_w;     /* host [nregion][nseg]                             */
  const float* col_w;      /* host [ncol] or NULL (== all ones)                */
  const float* cell_w;     /* DEVICE [nrow][ncol] (dense) or NULL              */
  int32_t zero_skip;
} wb2_weights;

/* ---- library / context ---------------------------------------------------- */
int wb2_version(void);
const char* wb2_last_error(void);
/* 1 when the library was built with the CUDA kernels (always, for this .so)   */
int wb2_has_cuda(void);
/* number of kernels this context has launched since creation (bench evidence) */
int64_t wb2_launch_count(const wb2_ctx* ctx);

int wb2_create(int device, wb2_ctx** out);
int wb2_destroy(wb2_ctx* ctx);
/* use an external cudaStream_t (e.g. torch's current stream); NULL = own stream */
int wb2_set_stream(wb2_ctx* ctx, void* cuda_stream);
void* wb2_get_stream(wb2_ctx* ctx);
int wb2_synchronize(wb2_ctx* ctx);
/* Ordering with a caller-owned stream WITHOUT blocking the host (the context's
 * own stream is cudaStreamNonBlocking, so nothing orders it implicitly against
 * e.g. torch's current stream):
 *   wb2_wait_stream : work enqueued on the context AFTER this call waits for
 *                     everything already enqueued on `cuda_stream` (inputs
 *                     still being produced there);
 *   wb2_stream_wait : work enqueued on `cuda_stream` AFTER this call waits for
 *                     everything already enqueued on the context (results).
 * No-ops when `cuda_stream` is the context's stream (wb2_set_stream).  NULL is
 * the legacy default stream.                                                  */
int wb2_wait_stream(wb2_ctx* ctx, void* cuda_stream);
int wb2_stream_wait(wb2_ctx* ctx, void* cuda_stream);

int wb2_malloc(wb2_ctx* ctx, size_t bytes, void** dev_ptr);
int wb2_free(wb2_ctx* ctx, void* dev_ptr);
int wb2_host_alloc(wb2_ctx* ctx, size_t bytes, void** host_ptr); /* pinned */
int wb2_host_free(wb2_ctx* ctx, void* host_ptr);
int wb2_memcpy_h2d(wb2_ctx* ctx, void* dev_dst, const void* host_src, size_t bytes);
int wb2_memcpy_d2h(wb2_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes);
int wb2_memset(wb2_ctx* ctx, void* dev_ptr, int value, size_t bytes);

/* ---- K1: deterministic metrics ---------------------------------------------
 * Replaces the arithmetic of  MSE / RMSESqrtBeforeTimeAvg / MAE / Bias / ACC
 * .compute_chunk (weatherbench2/metrics.py:251-269, 283-301, 323-330, 352-359,
 * 387-414) and of _spatial_average (metrics.py:141-163) for ALL regions of an
 * eval config in one pass over the data (evaluation.py:408-435 loops
 * metric x region and re-reads the chunk every time).
 *   f, t, c      device base pointers (c may be NULL: no climatology -> stats
 *                3,4,5,7,8,9 are written as 0)
 *   dtype        WB2_F32 / WB2_F64 element type of f, t and c
 *   off_f/t/c    host [nfield] element offsets of each field's slab
 *   skipna       0: NaN propagates into the sums (stats 6..9 then hold the
 *                plain weight sums);  1: NaN cells are skipped per statistic
 *   out          device [nfield][nregion][WB2_DET_NSTAT] float64
 */
int wb2_det_metrics(wb2_ctx* ctx, const void* f, const void* t, const void* c,
                    int dtype, int64_t nfield, const int64_t* off_f,
                    const int64_t* off_t, const int64_t* off_c,
                    const wb2_weights* w, int skipna, double* out);

/* Wind-vector variant: stat[0] = sum W*(du^2 + dv^2), stat[6] its weight sum
 * (WindVectorMSE.compute_chunk, metrics.py:189-202); other stats are 0.       */
int wb2_det_metrics_vector(wb2_ctx* ctx, const void* fu, const void* fv,
                           const void* tu, const void* tv, int dtype,
                           int64_t nfield, const int64_t* off_fu,
                           const int64_t* off_fv, const int64_t* off_tu,
                           const int64_t* off_tv, const wb2_weights* w,
                           int skipna, double* out);

/* Same as wb2_det_metrics but f, t, c are HOST base pointers and `out` is a
 * host array: the library streams the slabs through double-buffered device
 * staging (H2D on a copy stream overlapped with the kernel) and returns after
 * the result is in `out`.  This is the end-to-end entry the Python operators
 * use for NumPy inputs.  cell_w in `w` must still be a device pointer.        */
int wb2_det_metrics_host(wb2_ctx* ctx, const void* f, const void* t,
                         const void* c, int dtype, int64_t nfield,
                         const int64_t* off_f, const int64_t* off_t,
                         const int64_t* off_c, const wb2_weights* w, int skipna,
                         double* out_host);

/* ---- host-streaming entries and the slab cache ----------------------------------
 * Every *_host entry takes HOST base pointers (pinned memory reaches the PCIe
 * rate; pageable memory works at the driver's staging rate), streams the 2-D
 * slabs through two device staging buffers with the copies overlapped against
 * the kernels, and returns after the result is in host memory.
 *
 * wb2_set_slab_cache(ctx, bytes) reserves `bytes` of HBM as an LRU cache of
 * host slabs keyed by host address: the operands that repeat from chunk to
 * chunk -- the truth slabs of truth.sel(time=valid_time) (evaluation.py:475)
 * and the climatology slabs of the day-of-year lookup (metrics.py:398-404) --
 * then cross PCIe once instead of once per chunk (the forecast is never
 * cached).  The caller must not modify cached host arrays while the cache is
 * enabled; bytes = 0 (the default) disables it and frees the arena.
 * wb2_transfer_stats: which = 0 H2D bytes, 1 D2H bytes, 2 cache hits,
 * 3 cache misses of the *_host entries since the last reset.                     */
int wb2_set_slab_cache(wb2_ctx* ctx, size_t bytes);
int64_t wb2_transfer_stats(const wb2_ctx* ctx, int which);
int wb2_reset_transfer_stats(wb2_ctx* ctx);

/* wb2_ens_metrics on host buffers (x, t host base pointers, out host
 * [nfield][nregion][WB2_ENS_NSTAT]); truth slabs go through the slab cache.      */
int wb2_ens_metrics_host(wb2_ctx* ctx, const void* x, const void* t, int dtype,
                         int32_t nmember, int64_t member_stride, int64_t nfield,
                         const int64_t* off_x, const int64_t* off_t,
                         const wb2_weights* w, int skipna, double* out_host);

/* ---- K2: ensemble metrics --------------------------------------------------
 * Replaces CRPS / CRPSSkill / CRPSSpread, EnsembleMeanMSE / RMSE,
 * EnsembleVariance / Stddev, DebiasedEnsembleMeanMSE .compute_chunk
 * (metrics.py:657-715, 1185-1241, 1293-1363) incl. _rankdata (:836-846):
 * all five point-wise quantities from ONE read of the M members.
 *   x            device base pointer of the ensemble forecast
 *   off_x        host [nfield] offset of member 0 of each field
 *   member_stride elements between consecutive members of the same cell
 *   t, off_t     truth slab per field
 *   out          device [nfield][nregion][WB2_ENS_NSTAT] float64
 * 1..64 members: register sorting networks (the fast path); 65..1551 members
 * (metrics_test.py uses 100 and 1000): rank by counting in shared memory;
 * more: WB2_EUNSUPPORTED.
 */
int wb2_ens_metrics(wb2_ctx* ctx, const void* x, const void* t, int dtype,
                    int32_t nmember, int64_t member_stride, int64_t nfield,
                    const int64_t* off_x, const int64_t* off_t,
                    const wb2_weights* w, int skipna, double* out);

/* ---- K3: energy score -------------------------------------------------------
 * Replaces the per-member weighted L2 norms of EnergyScoreSkill / Spread
 * (metrics.py:1471-1517): every member is read once.
 *   out   device [nfield][nregion][4][nmember] float64:
 *         [0][m] sum W (x_m - t)^2         [1][m] sum W (x_m - x_{m+1})^2 (m < M-1)
 *         [2][m], [3][m] the matching weight sums
 * NaN propagates (skipna = False semantics); 1..64 members.                     */
int wb2_energy_score(wb2_ctx* ctx, const void* x, const void* t, int dtype,
                     int32_t nmember, int64_t member_stride, int64_t nfield,
                     const int64_t* off_x, const int64_t* off_t,
                     const wb2_weights* w, double* out);

/* ---- K6: map-output ("Spatial*") metrics with the time mean fused in ---------
 * Replaces SpatialBias / SpatialMSE / SpatialMAE .compute_chunk
 * (metrics.py:304-374) and the `.mean(time, skipna)` of Metric.compute
 * (metrics.py:117-138): out[j] = mean_g stat(f[j*ngroup+g], t[j*ngroup+g]) per
 * grid cell.  ngroup == 1 gives the per-time maps of compute_chunk unchanged.
 *   stat        WB2_MAP_BIAS f - t | WB2_MAP_MSE (f - t)^2 | WB2_MAP_MAE |f - t|
 *   off_f/off_t host [nout * ngroup] element offsets of the slabs (group-major)
 *   nrow, ncol, row_stride   slab geometry (col contiguous)
 *   skipna      mean over the non-NaN terms (NaN when there is none)
 *   out         device [nout][nrow][ncol], same element type as the inputs    */
enum { WB2_MAP_BIAS = 0, WB2_MAP_MSE = 1, WB2_MAP_MAE = 2 };
int wb2_det_maps(wb2_ctx* ctx, const void* f, const void* t, int dtype, int stat,
                 int64_t nout, int32_t ngroup, const int64_t* off_f,
                 const int64_t* off_t, int32_t nrow, int32_t ncol,
                 int64_t row_stride, int skipna, void* out);

/* Ensemble variant: SpatialCRPS / SpatialCRPSSkill / SpatialCRPSSpread
 * (metrics.py:718-772), SpatialEnsembleVariance (:1244-1266),
 * SpatialEnsembleMeanMSE / DebiasedSpatialEnsembleMeanMSE (:1366-1399) and the
 * time mean of EnsembleMetric.compute (:598-607).  One read of the members
 * yields every selected map.
 *   stat_mask   bit i selects point-wise statistic i of WB2_ENS_NSTAT's list
 *               (0 skill, 1 spread, 2 (t - xbar)^2, 3 variance, 4 debiased MSE);
 *               bit 5 = point-wise CRPS, skill - spread / 2 (metrics.py:729-739)
 *   out         device [popcount(stat_mask)][nout][nrow][ncol] float32, in
 *               increasing bit order; 1..64 members                            */
int wb2_ens_maps(wb2_ctx* ctx, const void* x, const void* t, int dtype,
                 int32_t nmember, int64_t member_stride, int64_t nout,
                 int32_t ngroup, const int64_t* off_x, const int64_t* off_t,
                 int32_t nrow, int32_t ncol, int64_t row_stride,
                 int32_t stat_mask, int skipna, float* out);

/* ---- K9: SEEPS maps ---------------------------------------------------------------
 * Replaces SpatialSEEPS.compute_chunk (metrics.py:417-513) and, for ngroup > 1,
 * the time mean of Metric.compute: dry / light / heavy categories of forecast
 * and truth precipitation, the 3 x 3 scoring matrix of the cell's
 * climatological dry fraction p1, NaN where p1 is outside (min_p1, max_p1).
 * SEEPS (the spatial mean with skipna, :516-528) is wb2_det_metrics on the maps.
 *   wet          device base of the climatological wet-threshold slabs;
 *   off_wet_f/t  host [nout * ngroup]: the slab for the forecast's / truth's
 *                valid time (day-of-year / hour lookup folded into the table)
 *   p1           device [nrow][ncol] float32 mean dry fraction
 *   out          device [nout][nrow][ncol] float32                               */
int wb2_seeps_maps(wb2_ctx* ctx, const float* f, const float* t, const float* wet,
                   const float* p1, int64_t nout, int32_t ngroup,
                   const int64_t* off_f, const int64_t* off_t,
                   const int64_t* off_wet_f, const int64_t* off_wet_t,
                   int32_t nrow, int32_t ncol, int64_t row_stride,
                   int64_t wet_row_stride, float dry_threshold, float min_p1,
                   float max_p1, int skipna, float* out);

/* ---- derived variables -------------------------------------------------------------
 * wb2_wind_speed replaces WindSpeed.compute (derived_variables.py:77-99):
 * out[i] = sqrt(u[i]^2 + v[i]^2) in float32, bit-identical to NumPy; device
 * pointers, n elements.                                                          */
int wb2_wind_speed(wb2_ctx* ctx, const float* u, const float* v, float* out,
                   int64_t n);

/* wb2_ens_mean replaces the ensemble-mean pipeline of
 * scripts/compute_ensemble_mean.py:110-141 (xbeam.Mean(realization, skipna)):
 * out[field][cell] = mean_m x[off_x[field] + m * member_stride + cell], float32;
 * skipna: mean over the non-NaN members (NaN when there is none).  Device
 * pointers; `slab` = cells per field (contiguous); out [nfield][slab].  The mean
 * stays in HBM for K1 (ensemble-mean RMSE / ACC) -- no host round trip.        */
int wb2_ens_mean(wb2_ctx* ctx, const float* x, int32_t nmember,
                 int64_t member_stride, int64_t nfield, const int64_t* off_x,
                 int64_t slab, int skipna, float* out);

/* wb2_spectrum_interp replaces interpolate_spectral_frequencies
 * (derived_variables.py:629-682) on the output of wb2_zonal_spectrum: row `lat`
 * of a spectrum lives on the increasing frequencies freq_table[lat][k] (host
 * [nrow][nk] float64, the `frequency` coordinate of the reference's result);
 * every row is interpolated linearly to the common `freqs` (host [nfreq],
 * float64), NaN outside the row's range like xarray's .interp.
 *   spec device [nfield][nrow][nk] float32;  out device [nfield][nrow][nfreq]    */
int wb2_spectrum_interp(wb2_ctx* ctx, const float* spec, int64_t nfield,
                        int32_t nrow, int32_t nk, const double* freq_table,
                        int32_t nfreq, const double* freqs, float* out);

/* ---- K10: rank histogram ---------------------------------------------------------
 * Replaces RankHistogram.compute_chunk (metrics.py:1894-2042) and the time mean
 * of EnsembleMetric.compute: rank of the truth among the members (NaN last),
 * binned into nbins (nbins divides nmember + 1), one-hot per grid point or,
 * for ngroup > 1, the mean of the one-hots over the group.  random_ties != 0
 * places the truth uniformly among members exactly equal to it (the effect of
 * the reference's tie-breaking noise) using a counter-based hash of `seed`.
 *   out   device [nout][nrow][ncol][nbins] float32                                */
int wb2_rank_histogram(wb2_ctx* ctx, const float* x, const float* t,
                       int32_t nmember, int64_t member_stride, int64_t nout,
                       int32_t ngroup, const int64_t* off_x, const int64_t* off_t,
                       int32_t nrow, int32_t ncol, int64_t row_stride,
                       int32_t nbins, int32_t random_ties, uint64_t seed,
                       float* out);

/* ---- K7: threshold ("binary event") and Gaussian-forecast metrics -------------
 * wb2_ens_threshold_metrics replaces EnsembleBrierScore /
 * DebiasedEnsembleBrierScore (metrics.py:1523-1710), EnsembleIgnoranceScore
 * (:1713-1790) and EnsembleRPS (:1793-1891) .compute_chunk: one read of the
 * members yields all four point-wise scores for every threshold.
 * Thresholds, two forms:
 *   thr_b == NULL  thr_a is a float32 threshold field per (threshold, field):
 *                  off_a host [nthreshold][nfield]  (QuantileThreshold,
 *                  thresholds.py:118-149)
 *   thr_b != NULL  thr_a / thr_b are the climatological mean / std slabs,
 *                  off_a / off_b host [nfield], z host [nthreshold] =
 *                  norm.ppf(quantile): thr = mean + z * std in float64
 *                  (GaussianQuantileThreshold, thresholds.py:152-185)
 *   out   device [nfield][nthreshold][nregion][8] float64:
 *         [0] sum W*brier  [1] sum W*debiased brier  [2] sum W*ignorance
 *         [3] sum W*rps part   [4..7] the matching weight sums
 * Any number of members (streamed, not held in registers).                      */
int wb2_ens_threshold_metrics(wb2_ctx* ctx, const void* x, const void* t,
                              int dtype, int32_t nmember, int64_t member_stride,
                              int64_t nfield, const int64_t* off_x,
                              const int64_t* off_t, int32_t nthreshold,
                              const void* thr_a, const int64_t* off_a,
                              const void* thr_b, const int64_t* off_b,
                              const double* z, const wb2_weights* w, int skipna,
                              double* out);

/* Map-output variant: SpatialEnsembleBrierScore (metrics.py:1615-1637),
 * SpatialDebiasedEnsembleBrierScore (:1701-1710), SpatialEnsembleIgnoranceScore
 * (:1768-1790), SpatialEnsembleRPS (:1868-1891) and the time mean of
 * EnsembleMetric.compute (:598-607): fields grouped like wb2_det_maps
 * (off_* host [nout * ngroup], field thresholds [nthreshold][nout * ngroup]).
 *   stat  0 brier | 1 debiased brier | 2 ignorance | 3 rps part
 *   out   device [nthreshold][nout][nrow][ncol] float32                          */
int wb2_ens_threshold_maps(wb2_ctx* ctx, const void* x, const void* t, int dtype,
                           int32_t nmember, int64_t member_stride, int64_t nout,
                           int32_t ngroup, const int64_t* off_x,
                           const int64_t* off_t, int32_t nthreshold,
                           const void* thr_a, const int64_t* off_a,
                           const void* thr_b, const int64_t* off_b,
                           const double* z, int32_t nrow, int32_t ncol,
                           int64_t row_stride, int32_t stat, int skipna,
                           float* out);

/* wb2_gaussian_metrics replaces GaussianCRPS / GaussianVariance
 * (metrics.py:849-937) when nthreshold == 0:
 *         out [nfield][1][nregion][8]: [0] sum W*crps [1] sum W*std^2, [4],[5] weights
 * and GaussianBrierScore / GaussianIgnoranceScore / GaussianRPS (:963-1158)
 * when nthreshold > 0 (thresholds as above):
 *         out [nfield][nthreshold][nregion][8]: [0] brier [2] ignorance [3] rps part
 * Point-wise math in float64 (the reference calls scipy.stats.norm in float64). */
int wb2_gaussian_metrics(wb2_ctx* ctx, const void* mean, const void* std,
                         const void* t, int dtype, int64_t nfield,
                         const int64_t* off_mean, const int64_t* off_std,
                         const int64_t* off_t, int32_t nthreshold,
                         const void* thr_a, const int64_t* off_a,
                         const void* thr_b, const int64_t* off_b,
                         const double* z, const wb2_weights* w, int skipna,
                         double* out);

/* ---- K5: conservative regridding -------------------------------------------
 * Replaces ConservativeRegridder.regrid_array (= _nanmean,
 * weatherbench2/regridding.py:502-536).  The two weight matrices
 * (regridding.py:341-373, 462-499) are banded; they are passed in CSR form
 * built by the Python operator with the reference's own formulas.
 * Input slabs are (lon_src, lat_src) with lat contiguous, output slabs
 * (lon_tgt, lat_tgt) float32, exactly the reference's layout (regridding.py:190).
 * A NaN weight row (target cell not covered, regridding.py:367-371) is flagged
 * by a negative tap count and yields NaN.
 */
typedef struct {
  int32_t n_src;            /* source axis length                              */
  int32_t n_tgt;            /* target axis length                              */
  const int32_t* row_ptr;   /* host [n_tgt + 1] CSR offsets; row_ptr[i+1] <
                               row_ptr[i] never happens; uncovered rows have
                               zero taps and nan_row[i] = 1                    */
  const int32_t* col_idx;   /* host [nnz] source indices                       */
  const float* val;         /* host [nnz] weights (rows sum to 1)              */
  const uint8_t* nan_row;   /* host [n_tgt] 1 = output NaN                     */
} wb2_csr;

int wb2_regrid_conservative(wb2_ctx* ctx, const float* src, float* dst,
                            int64_t nfield, int64_t src_field_stride,
                            int64_t dst_field_stride, const wb2_csr* lon_w,
                            const wb2_csr* lat_w);

/* wb2_regrid_conservative on host buffers: groups of fields in, regridded
 * groups out (H2D, kernel and D2H of neighbouring groups overlap).             */
int wb2_regrid_conservative_host(wb2_ctx* ctx, const float* src, float* dst,
                                 int64_t nfield, int64_t src_field_stride,
                                 int64_t dst_field_stride, const wb2_csr* lon_w,
                                 const wb2_csr* lat_w);

/* ---- K8: nearest-neighbour and bilinear regridding -----------------------------
 * wb2_regrid_gather replaces NearestRegridder.regrid_array
 * (regridding.py:231-247): dst[k] = src[indices[k]], `indices` host [ntarget]
 * into the raveled (lon, lat) source slab (BallTree / haversine nearest
 * neighbours, computed by the binder like regridding.py:212-228).
 * wb2_regrid_bilinear replaces BilinearRegridder.regrid_array (:256-294): per
 * target longitude a / latitude c two source taps and the fraction
 * delta / dx of jnp.interp (host arrays; tap -1 = NaN outside the source);
 * latitude is interpolated first, then longitude, in float32.
 * Slabs are (lon, lat) with lat contiguous; strides in elements.               */
int wb2_regrid_gather(wb2_ctx* ctx, const float* src, float* dst, int64_t nfield,
                      int64_t src_field_stride, int64_t dst_field_stride,
                      int32_t nsource, int32_t ntarget, const int32_t* indices);
int wb2_regrid_bilinear(wb2_ctx* ctx, const float* src, float* dst,
                        int64_t nfield, int64_t src_field_stride,
                        int64_t dst_field_stride, int32_t nlon_s, int32_t nlat_s,
                        int32_t nlon_t, int32_t nlat_t, const int32_t* lon_i0,
                        const int32_t* lon_i1, const float* lon_t,
                        const int32_t* lat_i0, const int32_t* lat_i1,
                        const float* lat_t);

/* ---- K4: zonal energy spectrum ---------------------------------------------
 * Replaces ZonalEnergySpectrum.compute (weatherbench2/derived_variables.py:
 * 592-626): rfft(norm='forward') along longitude, |F_k|^2 * (1 if k==0 else 2),
 * times `scale[row]` (the latitude circumference, derived_variables.py:626).
 *   x      device [nfield][nrow][ncol] float32, ncol = number of longitudes
 *   scale  host [nrow] float64
 *   out    device float32; accumulate == 0: [nfield][nrow][ncol/2+1] written;
 *          accumulate == 1: out has [nfield_out][nrow][ncol/2+1] and field i
 *          is ADDED to slot  i % nfield_out  (device-side time-sum for the
 *          script's `xbeam.Mean(['time'])`,
 *          scripts/compute_zonal_energy_spectrum.py:234)
 */
int wb2_zonal_spectrum(wb2_ctx* ctx, const float* x, int64_t nfield,
                       int32_t nrow, int32_t ncol, const double* scale,
                       float* out, int32_t accumulate, int64_t nfield_out);

/* K4 with the weighted meridional reduction fused in (BASELINE north star: "a
 * shared-memory rFFT along longitude followed by a weighted meridional
 * reduction").  The reference keeps `latitude` (derived_variables.py:592-626);
 * its callers average the per-latitude spectra over latitude bands afterwards,
 * so the reduction is defined on the reference's own output:
 *     out[slot][k] = sum_{fields i of the slot} sum_row scale[row] * S_ref[i][row][k]
 * where S_ref is the spectrum WITHOUT the circumference factor, i.e. the caller
 * passes scale[row] = circumference(row) * w(row) with w the (normalised)
 * latitude weight of weatherbench2/metrics.py:40-60 (or a latitude-band mask
 * times it).  Fields are slot-minor like wb2_zonal_spectrum's accumulate mode
 * (field i belongs to slot i % nfield_out), so a time mean is the same call.
 *   out    device [nfield_out][ncol/2+1] float32, overwritten
 * The per-latitude spectrum is never written: 4 B read per cell, ~0 written. */
int wb2_zonal_spectrum_latsum(wb2_ctx* ctx, const float* x, int64_t nfield,
                              int32_t nrow, int32_t ncol, const double* scale,
                              float* out, int64_t nfield_out);

/* wb2_zonal_spectrum / wb2_zonal_spectrum_latsum on host buffers.  x host
 * [nfield][nrow][ncol] dense; out host, OVERWRITTEN: accumulate == 0:
 * [nfield][nrow][ncol/2+1]; accumulate != 0: the sum over the fields of each
 * slot, [nfield_out][nrow][ncol/2+1] (the accumulator stays in HBM, only the
 * sum comes back); latsum: [nfield_out][ncol/2+1].                             */
int wb2_zonal_spectrum_host(wb2_ctx* ctx, const float* x, int64_t nfield,
                            int32_t nrow, int32_t ncol, const double* scale,
                            float* out_host, int32_t accumulate,
                            int64_t nfield_out);
int wb2_zonal_spectrum_latsum_host(wb2_ctx* ctx, const float* x, int64_t nfield,
                                   int32_t nrow, int32_t ncol,
                                   const double* scale, float* out_host,
                                   int64_t nfield_out);

#ifdef __cplusplus
}
#endif
#endif /* WB2B200_H_ */
