/*
 * metada_cuda_c_api.h -- thin C ABI of the B200 (sm_100a) backend for METADA's ensemble Kalman
 * analysis step (LETKF / ETKF / EnKF).
 *
 * This is the drop-in boundary: the C++ host classes in metada_b200/host (CudaBackendTag traits,
 * CudaState, CudaObservation, CudaObsOperator, device Ensemble store, LETKF/ETKF/EnKF policies)
 * call only these entry points; so does the Python ctypes binding used by tests and bench.py.
 * Conventions follow the reference's own native bridges (opaque handles with create/destroy as in
 * backends/lorenz63/state/lorenz63/state_c_api.h:7-16; int return code, 0 = success, caller-owned
 * output buffers as in backends/common/obsoperator/WRFDAObsOperator_c_api.h:35-130).
 *
 *   - extern "C", plain pointers and sizes; no C++/torch types.
 *   - every call returns 0 on success, non-zero on failure; mdc_last_error(ctx) has the message.
 *   - one context per device; handles are not thread-safe; all work is queued on the context's
 *     stream; calls that fill HOST buffers synchronise before returning.
 *   - there is NO CPU fallback: without a CUDA device mdc_ctx_create fails.
 *
 * Reference paths below are relative to /root/reference/src.
 */
#ifndef METADA_CUDA_C_API_H
#define METADA_CUDA_C_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mdc_ctx mdc_ctx;
typedef struct mdc_ens mdc_ens;
typedef struct mdc_obs mdc_obs;

#define MDC_OK 0
#define MDC_ERR_INVALID 1
#define MDC_ERR_CUDA 2
#define MDC_ERR_NUMERIC 3
#define MDC_ERR_UNSUPPORTED 4

/* ---- context ---------------------------------------------------------------------------- */
int mdc_ctx_create(int device, mdc_ctx** out);
int mdc_ctx_destroy(mdc_ctx* ctx);
const char* mdc_last_error(const mdc_ctx* ctx);
int mdc_ctx_sync(mdc_ctx* ctx);
/* cudaStream_t of the context, for interop (e.g. torch.cuda.ExternalStream) */
void* mdc_ctx_stream(mdc_ctx* ctx);
/* CUDA-event stopwatch on the context's stream: start, ... work ..., stop -> elapsed ms */
int mdc_timer_start(mdc_ctx* ctx);
int mdc_timer_stop(mdc_ctx* ctx, float* ms);
/* number of kernels this library has launched on ctx since creation (bench.py gpu_launches) */
int64_t mdc_ctx_launch_count(const mdc_ctx* ctx);
int mdc_ctx_sm_count(const mdc_ctx* ctx);
/* the 16 raw device counters of the last analysis on ctx (development diagnostics: [8..15] hold per-phase clock
 * ticks when the column kernel is built with -DNSP_PROFILE, zeros otherwise) */
int mdc_ctx_last_stats(mdc_ctx* ctx, int64_t out[16]);
/* writes > L2-size bytes to evict the L2 between timed iterations */
int mdc_ctx_flush_l2(mdc_ctx* ctx);
/* plain device buffers (halo staging between observation stores on one device) */
int mdc_dev_malloc(mdc_ctx* ctx, int64_t bytes, void** out);
int mdc_dev_free(mdc_ctx* ctx, void* p);
/* kind 1: host -> device, 2: device -> host (synchronous on the context's stream) */
int mdc_dev_copy(mdc_ctx* ctx, void* dst, const void* src, int64_t bytes, int kind);

/* ---- ensemble store: replaces framework/adapters/Ensemble.hpp:42-194 (vector<State>) ------
 * Device layout: X[col][lev][member], col = y*nx + x  (one contiguous nz*k block per column).
 * Host layout of one member (State::getDataPtr<double>, State.hpp:229-242; SimpleState.hpp:73-80):
 *   [lev][y][x].
 * (nx, ny) is the LOCAL grid including any read-only halo; set_domain places it in the global
 * grid for multi-GPU column sharding (default: the local grid is the whole domain). */
int mdc_ens_create(mdc_ctx* ctx, int nx, int ny, int nz, int k, mdc_ens** out);
int mdc_ens_destroy(mdc_ens* ens);
int mdc_ens_set_domain(mdc_ens* ens, int gx0, int gy0, int gnx, int gny, int own_nx, int own_ny);
/* reuse one allocation for slabs of different heights: ny <= the ny given to mdc_ens_create */
int mdc_ens_set_rows(mdc_ens* ens, int ny);
int mdc_ens_upload_member(mdc_ens* ens, int m, const double* host);
int mdc_ens_download_member(mdc_ens* ens, int m, double* host);
/* batched variants (coalesced transposes): members m0..m0+count-1, one host pointer each */
int mdc_ens_upload_members(mdc_ens* ens, int m0, int count, const double* const* hosts);
int mdc_ens_download_members(mdc_ens* ens, int m0, int count, double* const* hosts);
/* Row-slab variants for streaming a large host ensemble through the device in pieces: hosts[m]
 * points at the FULL member array [lev][host_ny][nx]; upload reads its rows host_y0 .. host_y0+ny-1
 * (ny = this store's local rows), download writes local rows 0 .. nrows-1 to host rows host_y0 ..
 * (nrows < ny leaves a read-only halo row untouched).  One strided copy per member and batch. */
int mdc_ens_upload_members_rows(mdc_ens* ens, int m0, int count, const double* const* hosts,
                                int host_ny, int host_y0);
int mdc_ens_download_members_rows(mdc_ens* ens, int m0, int count, double* const* hosts,
                                  int host_ny, int host_y0, int nrows);
/* seeded synthetic ensemble generated on the device (bit-identical to
 * metada_b200.synthetic.member on the host); coordinates are global. */
int mdc_ens_fill_synthetic(mdc_ens* ens, uint64_t seed);
/* Ensemble::RecomputeMean (Ensemble.hpp:105-114) on the device; host_mean [lev][y][x] or NULL */
int mdc_ens_mean(mdc_ens* ens, double* host_mean);
/* sum over members/levels/columns of X and X^2 -- cheap whole-state checksum for big runs */
int mdc_ens_checksum(mdc_ens* ens, double* sum, double* sumsq);
double* mdc_ens_devptr(mdc_ens* ens);
int64_t mdc_ens_bytes(const mdc_ens* ens);

/* ---- observations: replaces backends/common/observation/GridObservation.hpp (AoS) ---------
 * SoA on the device. GRID coordinates (Location(int,int,int), Location.hpp:66-67) are GLOBAL.
 * err is the standard deviation; variance = err*err (GridObservation.hpp:239-252); invalid obs
 * have H(x) = 0 (IdentityObsOperator.hpp:165-168) and infinite variance.
 * gid: global observation ids (NULL = 0..P-1); they define the summation order. */
int mdc_obs_create(mdc_ctx* ctx, int64_t P, const int32_t* x, const int32_t* y, const int32_t* z,
                   const double* value, const double* err, const uint8_t* valid,
                   const int64_t* gid, mdc_obs** out);
/* replace the contents of an existing store (buffers are reused / grown; halo rows, H(x) results
 * and the index are dropped) */
int mdc_obs_assign(mdc_obs* obs, int64_t P, const int32_t* x, const int32_t* y, const int32_t* z,
                   const double* value, const double* err, const uint8_t* valid, const int64_t* gid);
int mdc_obs_destroy(mdc_obs* obs);
int64_t mdc_obs_size(const mdc_obs* obs);       /* own + halo rows */

/* ---- geographic observations and multi-variable states (the WRF-shaped case) ------------------
 * Geography of the grid: latitude / longitude in degrees of every column, [ny][nx] (the 2-D coordinate arrays of
 * WRFGeometry::unstaggered_info(), read by IdentityObsOperator.hpp:488-509 and WRFGeometryIterator.hpp:144), and the
 * geometry's vertical coordinate (nlev values, or nlev = 0 / NULL: every observation sits on level 0, :515-526). */
int mdc_ens_set_geography(mdc_ens* ens, const double* lat, const double* lon, int nlev,
                          const double* vertical_coords);
/* Variables of the state: a member is [var][lev][y][x] (WRFState.hpp:866-877), i.e. nz = sum of var_nlev; 3-D
 * variables share the geometry's levels, 2-D variables have 1 level.  H reads an observation's own variable
 * (IdentityObsOperator.hpp:681-711); every variable of a column is updated with the column's transform; with vertical
 * localisation the distance is taken between levels inside their variables.  At most 16 variables. */
int mdc_ens_set_variables(mdc_ens* ens, int nvar, const int32_t* var_nlev);
/* state variable each observation observes (index into the ensemble's variables); NULL = variable 0 */
int mdc_obs_set_variables(mdc_obs* obs, const int32_t* var);
/* Observations with GEOGRAPHIC locations (Location(lat, lon, level, GEOGRAPHIC), Location.hpp:82-84): the local
 * selection is Location::distance_to = haversine kilometres on a sphere of radius 6371 km (Location.hpp:213-217,
 * 325, 349-357), so mdc_letkf_params.radius (and loc_scale) are kilometres; mdc_obs_locate finds each observation's
 * nearest grid point and level exactly as IdentityObsOperator::convertGeographicToGrid (:484-530) -- first minimum of
 * the Euclidean distance in degrees -- and H then interpolates at those integer coordinates (:241-248).
 * mdc_hx_idw4 / mdc_letkf_analyse locate on demand.  Regional domains only: the selection circles must not reach a
 * pole or wrap the whole longitude circle (MDC_ERR_UNSUPPORTED otherwise).  All three modes.  Domain decomposition:
 * locate on a store that covers the whole grid, give every rank's store its window of that geography
 * (mdc_ens_set_geography_from) and exchange the halo rows by box (mdc_obs_pack_rows_geo).  Staggered (U / V) variables: give the staggered grid its own ensemble store and geography,
 * run mdc_hx_idw4(mass_grid_ens, obs) and then mdc_letkf_analyse(staggered_ens, obs, ...) -- the analysis uses the
 * Y' already in the store and the columns / coordinates of the ensemble it is given. */
int mdc_obs_create_geographic(mdc_ctx* ctx, int64_t P, const double* lat, const double* lon,
                              const double* level, const double* value, const double* err,
                              const uint8_t* valid, const int64_t* gid, mdc_obs** out);
int mdc_obs_locate(mdc_obs* obs, mdc_ens* ens);
/* Geography of a DECOMPOSED store (mdc_ens_set_domain) = its window of `global`, a store that covers the whole grid
 * and has its geography set (one level and one member are enough: it only serves mdc_obs_locate and this call).  The
 * window inherits the global frame -- unwrap centre and extents --, so the lat / lon lattice of the bucket index, and
 * with it the order in which a column meets its candidates, is the same for every decomposition: the analysis is
 * bit-identical to the single-store one. */
int mdc_ens_set_geography_from(mdc_ens* ens, const mdc_ens* global);
/* unwrap centre of the longitudes, extents of the unwrapped longitude offsets u = (lon - lon_c) - 360 rint(..) and of
 * the latitudes over the store's columns (any output may be NULL) */
int mdc_ens_geography_frame(const mdc_ens* ens, double* lon_c, double* umin, double* umax, double* latmin,
                            double* latmax);
/* grid coordinates of the observations (as given, or as located); any output may be NULL */
int mdc_obs_download_grid_coords(mdc_obs* obs, int32_t* x, int32_t* y, int32_t* z);

/* ---- H(x), Y' and d: replaces k calls of ObsOperator::apply (ObsOperator.hpp:255-259 ->
 * IdentityObsOperator.hpp:154-180, 594-676) + LETKF.hpp:209-211 / ETKF.hpp:135-141 ---------- */
int mdc_hx_idw4(mdc_ens* ens, mdc_obs* obs);
/* any output may be NULL; Y, Yp are [P][k] row-major */
int mdc_hx_download(mdc_obs* obs, double* Y, double* ybar, double* Yp, double* d);

/* ---- multi-GPU observation halo (column sharding) ------------------------------------------
 * Rows are (k + 8) doubles: Y'[k], d, value, err, valid, x, y, z, gid -- (k + 12) for geographic and per-variable
 * stores, which add lat, lon, level, variable (mdc_obs_row_doubles tells).  pack selects own obs with
 * ylo <= y < yhi into a DEVICE buffer; append adds received rows as halo obs.  pack_rows_geo selects by a box in
 * the frame of the geography instead (lat_lo <= lat <= lat_hi, u_lo <= unwrapped lon - lon_c <= u_hi): the
 * destination's columns' bounding box widened by the radius -- a cover, the haversine selection decides. */
int mdc_obs_pack_rows(mdc_obs* obs, int ylo, int yhi, double* dev_rows, int64_t cap, int64_t* n);
int mdc_obs_pack_rows_geo(mdc_obs* obs, double lat_lo, double lat_hi, double u_lo, double u_hi, double lon_c,
                          double* dev_rows, int64_t cap, int64_t* n);
int mdc_obs_append_rows(mdc_obs* obs, const double* dev_rows, int64_t n);
int mdc_obs_row_doubles(const mdc_obs* obs);

/* ---- bucketed spatial index: replaces the O(P) scan of LETKF.hpp:159-165 -------------------
 * cell <= 0 picks ceil(radius). Selection is bit-identical to
 * Location::distance_to(...) <= radius (Location.hpp:204-211). */
int mdc_obs_index_build(mdc_obs* obs, int cell);
int mdc_obs_index_query_counts(mdc_obs* obs, mdc_ens* ens, double radius, int32_t* host_counts);
/* lists[c*cap .. ) = global obs ids of column cols[c] in kernel order; counts[c] = full count */
int mdc_obs_index_query_lists(mdc_obs* obs, mdc_ens* ens, double radius, const int64_t* cols,
                              int64_t ncols, int32_t cap, int64_t* host_lists,
                              int32_t* host_counts);

/* ---- LETKF: replaces LETKF<Tag>::Analyse / updateGridPoint (LETKF.hpp:63-119, 152-243) ----- */
enum { MDC_MODE_REF_COMPAT = 0, MDC_MODE_REF_ETKF = 1, MDC_MODE_CANONICAL = 2 };
/* R-localisation weight rho(d) of an observation at distance d (CANONICAL mode; the selection cutoff d <= radius
 * always applies).  GASPARI_COHN: Gaspari & Cohn 1999 eq. 4.10 with support = radius.  The other three are the
 * reference's localisation functions (LWEnKF.hpp:597-635) of d / L, L = loc_scale (<= 0: L = radius):
 * GAUSSIAN exp(-(d/L)^2 / 2), EXPONENTIAL exp(-d/L), REF_GASPARI_COHN the reference's own two-piece polynomial
 * (LWEnKF.hpp:624-635) -- NOT the Gaspari-Cohn taper (it is 4 at d = 0 and discontinuous at d = L); kept only so
 * that the reference's choice can be reproduced. */
enum { MDC_LOC_CUTOFF = 0, MDC_LOC_GASPARI_COHN = 1, MDC_LOC_GAUSSIAN = 2, MDC_LOC_EXPONENTIAL = 3,
       MDC_LOC_REF_GASPARI_COHN = 4 };
/* AUTO: Newton-Schulz (GEMM-only symmetric square root on FP64 DMMA) for 24 <= k <= 128, else Jacobi.
 * JACOBI: one-block-per-column one-sided Jacobi eigen-decomposition with warp-shuffle reductions.
 * NEWTON_SCHULZ: packed symmetric tiles (two columns per SM for k <= 80); transforms whose
 *   conditioning bound exceeds mdc_letkf_params.kappa_max are redone by NEWTON_SCHULZ_FULL (k <= 80) or JACOBI (k > 80).
 * NEWTON_SCHULZ_FULL: every product computed in full, any conditioning, 24 <= k <= 80.
 * All give the same (unique) symmetric square-root transform to rounding. */
enum { MDC_SOLVER_AUTO = 0, MDC_SOLVER_JACOBI = 1, MDC_SOLVER_NEWTON_SCHULZ = 2, MDC_SOLVER_NEWTON_SCHULZ_FULL = 3 };

typedef struct {
  double radius;      /* horizontal selection radius (inclusive) = Gaspari-Cohn support       */
  double radius_v;    /* vertical radius in levels; <= 0: none (one transform per column)     */
  double inflation;
  int mode;           /* MDC_MODE_*                                                           */
  int loc;            /* MDC_LOC_* (CANONICAL)                                                */
  int use_R;          /* CANONICAL: 1 -> R = diag(err^2), 0 -> R = I                          */
  int max_sweeps;     /* Jacobi sweep cap (<= 0: 40)                                          */
  double jacobi_tol;  /* stop when a sweep's max |g_p.g_q|/(|g_p||g_q|) < tol (<= 0: 1e-11)    */
  int solver;         /* CANONICAL: how A^{-1/2} is formed -- MDC_SOLVER_*                    */
  int sm_reserve;     /* SMs the persistent column kernel leaves free for concurrent streams
                         (member transposes of a streamed pipeline); 0 = use every SM          */
  double loc_scale;   /* length scale L of MDC_LOC_GAUSSIAN / EXPONENTIAL / REF_GASPARI_COHN; <= 0: radius.
                         The vertical scale is radius_v * L / radius.                          */
  double kappa_max;   /* NEWTON_SCHULZ: largest rigorous condition bound lambda_max / lambda_min of a transform's
                         k x k matrix the packed symmetric kernel keeps (the others are redone, see above);
                         <= 0: 1e5 (the analysis agrees with the eigen-decomposition to ~5e-12 there: the mean update of
                         ill-conditioned transforms is iteratively refined), at most 3e5 (~1e-11)              */
} mdc_letkf_params;

typedef struct {
  float ms_hx, ms_index, ms_columns, ms_total;
  int64_t columns;        /* analysed columns                                                 */
  int64_t sum_local_obs;  /* sum over columns of p_loc                                        */
  int32_t max_local_obs;
  int32_t max_sweeps;     /* max Jacobi sweeps (or Newton-Schulz iterations) used by a column */
  int64_t sum_sweeps;
  int32_t numeric_failures;
  int32_t redo_transforms; /* transforms the packed Newton-Schulz kernel handed to its fallback */
  int64_t small_transforms; /* transforms done in observation space (p_loc <= 24 and 2 p_loc <= k)  */
} mdc_letkf_stats;

int mdc_letkf_analyse(mdc_ens* ens, mdc_obs* obs, const mdc_letkf_params* params,
                      mdc_letkf_stats* stats);
/* transform W (k x k row-major; REF_COMPAT: s_i in row 0) of column `col` at level 0, for tests */
int mdc_letkf_column_transform(mdc_ens* ens, mdc_obs* obs, const mdc_letkf_params* params,
                               int64_t col, double* host_W);

/* ---- global ETKF: replaces ETKF<Tag>::Analyse (ETKF.hpp:100-179) --------------------------- */
int mdc_etkf_analyse(mdc_ens* ens, mdc_obs* obs, double inflation);

/* ---- global stochastic EnKF: replaces EnKF<Tag>::Analyse (EnKF.hpp:139-256) ---------------- */
typedef struct {
  double innovation_norm, background_spread, analysis_spread;
  double max_kalman_gain, min_kalman_gain, condition_number;
} mdc_enkf_diag;
/* Z: host [P][k] standard-normal draws (obs_pert = sqrt(R_ii) Z, EnKF.hpp:340-361), or NULL to
 * draw them on the device from `seed`. want_gain_stats: stream max/min of K without storing it. */
int mdc_enkf_analyse(mdc_ens* ens, mdc_obs* obs, double inflation, const double* Z, uint64_t seed,
                     int want_gain_stats, mdc_enkf_diag* diag);

/* ---- locally weighted EnKF: replaces LWEnKF<Tag>::Analyse (LWEnKF.hpp:207-334) ------------------------------
 * As written in the reference: global and dense (S = (sum_m w_m y'_m y'_m^T) o L + R is P x P; K = X' Y'^T S^-1 /
 * (k - 1) is Schur-multiplied by a second localisation matrix; both use INDEX distances |i - j| / dim normalised by
 * loc_radius, LWEnKF.hpp:566-570, 589-594), member weights by `weighting` (:400-531), multiplicative inflation of
 * the perturbations (:641-660), perturbed observations (:665-685; Z as for mdc_enkf_analyse).  loc_fn: MDC_LOC_CUTOFF,
 * MDC_LOC_GAUSSIAN, MDC_LOC_EXPONENTIAL or MDC_LOC_REF_GASPARI_COHN (the reference's "gaspari_cohn").  K is never
 * stored.  Needs libcusolver at run time (LU of S, eigenvalues for cond(S)). */
enum { MDC_LW_UNIFORM = 0, MDC_LW_ADAPTIVE = 1, MDC_LW_INVERSE_VAR = 2, MDC_LW_LIKELIHOOD = 3 };
typedef struct {
  double innovation_norm, background_spread, analysis_spread, max_kalman_gain, min_kalman_gain, condition_number;
  double max_weight, min_weight, weight_variance;
} mdc_lwenkf_diag;
int mdc_lwenkf_analyse(mdc_ens* ens, mdc_obs* obs, double inflation, double loc_radius, int loc_fn, int weighting,
                       const double* Z, uint64_t seed, mdc_lwenkf_diag* diag);

/* ---- verification metrics against a truth state -------------------------------------------------
 * Replaces framework/algorithms/Metrics.hpp:74-290 (Metrics<T>::CalculateAll): ensemble mean and
 * spread (unbiased standard deviation) per state point, RMSE / bias / correlation of the mean
 * against the truth, CRPS (empirical-CDF form, :232-254) and the average spread.  `truth` is a
 * one-member ensemble on the same grid (mdc_ens_create(..., k = 1) + mdc_ens_upload_member).
 * host_spread: optional [nz][ny][nx] output of the spread field (member-major host order). */
typedef struct {
  double rmse, bias, correlation, crps, avg_spread;
} mdc_metrics;
int mdc_ens_metrics(mdc_ens* ens, mdc_ens* truth, mdc_metrics* out, double* host_spread);

/* ---- streaming and sharding runtime (csrc/mdc_runtime.cpp) -----------------------------------------------
 * The reference's LETKF<Tag>::Analyse (LETKF.hpp:63-119) walks one in-memory ensemble; its members are host vectors
 * reached through State::getDataPtr<double>() (State.hpp:229-242).  mdc_stream_analyse takes exactly those pointers
 * and streams the rows [row0, row1) of the grid through the device in slabs on three host threads (upload ||
 * H + halo between slabs + column analysis || download, one CUDA stream per slab in flight), in place, so an
 * ensemble larger than the device (BASELINE C5: 86 GB) is analysed by the same call; the result is bit-identical to
 * mdc_letkf_analyse on the whole grid.  MDC_MODE_CANONICAL and the REF modes; GRID observations.
 *
 * Column sharding: one process per GPU, rank r owns the rows of slab_bounds(gny, r, nranks) =
 * [gny r / nranks, gny (r + 1) / nranks); mdc_comm_init attaches an NCCL communicator (id from
 * mdc_comm_get_unique_id on rank 0, handed to the other ranks by the launcher) and mdc_stream_analyse then starts
 * with the observation halo exchange: H on the rank's edge strips, rows of the observations other ranks' columns
 * can reach packed per destination (Y'[k], d, value, err, valid, x, y, z, gid) and carried by one group of
 * ncclSend / ncclRecv over NVLink; no other communication.  Every rank passes the GLOBAL observation arrays (the
 * message sizes follow from them, no count exchange) and its own rows of the members. */
typedef struct mdc_stream mdc_stream;
typedef struct {
  int gnx, gny, nz, k;
  int row0, row1;     /* global rows analysed through this handle; the whole grid: 0, gny                    */
  int slab_rows;      /* rows per slab; <= 0: chosen so that the range has >= 24 slabs of 8..32 rows         */
  int slots;          /* slabs in flight (>= 3; raised when a slab is lower than the localisation reach)     */
  int sm_reserve;     /* SMs the column kernel leaves to the transposes of the neighbouring slabs (8)        */
  double radius;      /* largest horizontal radius mdc_stream_analyse will be given                          */
} mdc_stream_config;
int mdc_stream_create(int device, const mdc_stream_config* cfg, mdc_stream** out);
int mdc_stream_destroy(mdc_stream* s);
const char* mdc_stream_last_error(const mdc_stream* s);
int mdc_stream_slabs(const mdc_stream* s);
int mdc_stream_slots(const mdc_stream* s);
/* members: k host pointers (pinned for full PCIe speed) to [nz][host_ny][gnx] arrays whose first row is global row
 * host_row0; they must cover [row0, min(row1 + 1, gny)) (H reads one row above the range).  ox .. ovalid: the global
 * observation set (oz, ovalid may be NULL).  Updated in place; returns when every slab is back. */
int mdc_stream_analyse(mdc_stream* s, double* const* members, int host_row0, int host_ny, int64_t P,
                       const int32_t* ox, const int32_t* oy, const int32_t* oz, const double* oval,
                       const double* oerr, const uint8_t* ovalid, const mdc_letkf_params* params,
                       mdc_letkf_stats* out);
/* wall-clock milliseconds of the last call's halo prologue and slab pipeline, rows received from other ranks */
int mdc_stream_timings(const mdc_stream* s, double* edge_halo_ms, double* stream_ms, int64_t* halo_rows);
/* 128-byte NCCL unique id (rank 0) */
int mdc_comm_get_unique_id(void* id, int bytes);
int mdc_comm_init(mdc_stream* s, const void* id, int rank, int nranks);
/* after a sharded analysis on FULL host members [nz][gny][gnx] (host_row0 = 0): every rank's analysed rows are passed
 * to all the others, so that each process ends with the whole analysed ensemble (what a drop-in driver saves) */
int mdc_comm_allgather_rows(mdc_stream* s, double* const* members);
/* max over ranks of *value (device-side ncclAllReduce); single rank: unchanged */
int mdc_comm_max(mdc_stream* s, double* value);
/* Sharded analysis of GEOGRAPHIC observations through the same handle (created for this rank's slab, mdc_comm_init
 * done; one process: the whole grid).  Every rank passes the GLOBAL latitude / longitude arrays [gny][gnx], the
 * vertical coordinate, the variables' level counts (nvar = 0: one variable) and ALL observations (level, valid,
 * variable may be NULL); members[m] points at the full host member [nz][gny][gnx], of which the rank's rows are
 * read (+ one halo row) and overwritten with the analysis (mdc_comm_allgather_rows shares them afterwards).  The
 * result is bit-identical to mdc_letkf_analyse on one store with mdc_ens_set_geography.  radius in kilometres. */
int mdc_geo_sharded_analyse(mdc_stream* s, double* const* members, const double* glat, const double* glon, int nvc,
                            const double* vertical_coords, int nvar, const int32_t* var_nlev, int64_t P,
                            const double* olat, const double* olon, const double* olevel, const double* ovalue,
                            const double* oerr, const uint8_t* ovalid, const int32_t* ovar,
                            const mdc_letkf_params* params, mdc_letkf_stats* stats);

/* ---- microbenchmarks used for the roofline denominators (profiles/) ------------------------ */
int mdc_bench_fp64_fma(mdc_ctx* ctx, double* tflops);
int mdc_bench_fp64_dmma(mdc_ctx* ctx, double* tflops);
int mdc_bench_hbm_copy(mdc_ctx* ctx, double* gbs);

#ifdef __cplusplus
}
#endif
#endif
