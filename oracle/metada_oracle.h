/*
 * metada_oracle.h -- CPU restatement of the METADA ensemble Kalman analysis path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the timed CPU baseline.
 *
 * Parity pin: the formulas restated here are checked (tests/test_oracle_vs_ref.py,
 * tests/golden/) against outputs of the reference's own, unmodified LETKF.hpp /
 * ETKF.hpp / EnKF.hpp compiled in oracle/_ref against a minimal Eigen-API shim
 * (Eigen itself is not in this image; see oracle/README.md).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src).
 *
 * Layouts (the reference's own, member-major):
 *   ensemble  X[m][lev][y][x]  -> X[m*n + (lev*ny + y)*nx + x],  n = nx*ny*nz
 *   obs       SoA: ox,oy,oz int32 GRID coordinates, value, err (std-dev), valid
 *   Y         [obs][member] row-major (P x k)
 */
#ifndef METADA_ORACLE_H
#define METADA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_MODE_REF_COMPAT = 0, ORC_MODE_REF_ETKF = 1, ORC_MODE_CANONICAL = 2 };
enum { ORC_LOC_CUTOFF = 0, ORC_LOC_GASPARI_COHN = 1, ORC_LOC_GAUSSIAN = 2, ORC_LOC_EXPONENTIAL = 3,
       ORC_LOC_REF_GASPARI_COHN = 4 };
enum { ORC_SEM_SNAPSHOT = 0, ORC_SEM_AS_WRITTEN = 1 };

typedef struct {
  int nx, ny, nz, k;      /* grid and ensemble size                                   */
  int64_t P;              /* number of observations                                   */
  double radius;          /* horizontal selection radius, inclusive (<=)              */
  double radius_v;        /* vertical radius in levels; <=0 : no vertical localisation */
  double inflation;       /* multiplicative inflation (meaning depends on mode)       */
  int mode;               /* ORC_MODE_*                                               */
  int loc;                /* ORC_LOC_* (CANONICAL only; REF modes are cutoff)         */
  int use_R;              /* CANONICAL: 1 = R = diag(err^2), 0 = R = I                */
  int semantics;          /* ORC_SEM_*                                                */
  int nthreads;           /* OpenMP threads for SNAPSHOT (<=0: all)                   */
  double loc_scale;       /* length scale of the exp-type functions; <= 0: radius     */
} orc_letkf_params;

/* LWEnKF::computeLocalizationFunction (LWEnKF.hpp:597-621) and its computeGaspariCohnFunction
 * (:624-635) for loc = GAUSSIAN / EXPONENTIAL / REF_GASPARI_COHN with normalised distance
 * dist / scale; loc = GASPARI_COHN is the Gaspari-Cohn 1999 taper with support `support`. */
double orc_loc_weight(int loc, double dist, double support, double scale);

/* Metrics<double>::CalculateAll (framework/algorithms/Metrics.hpp:74-103): X is [k][n] (member
 * major), truth [n]; mean/spread [n] may be NULL.  out = {rmse, bias, correlation, crps, avg_spread}. */
void orc_metrics(const double* X, const double* truth, int64_t n, int k, double* mean, double* spread,
                 double out[5]);

/* Location::distance_to for two GRID locations (Location.hpp:204-211). */
double orc_distance_grid(int i1, int j1, int i2, int j2);

/* LETKF.hpp:159-165: indices (ascending) of obs with distance <= radius. Returns count. */
int64_t orc_select_local(int gx, int gy, int64_t P, const int32_t* ox, const int32_t* oy,
                         double radius, int32_t* idx_out);

/* Counts only, for every column (y-outer, x-inner) -- brute force O(G*P). */
void orc_select_counts(int nx, int ny, int64_t P, const int32_t* ox, const int32_t* oy,
                       double radius, int32_t* counts, int nthreads);

/* IdentityObsOperator::apply (IdentityObsOperator.hpp:154-180, 594-676) for one member. */
void orc_hx_idw4(const double* member, int nx, int ny, int nz, int64_t P, const int32_t* ox,
                 const int32_t* oy, const int32_t* oz, const uint8_t* valid, double* out);

/* Ensemble::RecomputeMean (Ensemble.hpp:105-114): zero, += members in order, *= 1.0/k. */
void orc_ensemble_mean(const double* X, int k, int64_t n, double* mean);

/* Y (P x k), ybar, Y' = Y - ybar, d = yo - ybar   (LETKF.hpp:209-211, ETKF.hpp:128-141). */
void orc_obs_space(const double* X, int nx, int ny, int nz, int k, int64_t P, const int32_t* ox,
                   const int32_t* oy, const int32_t* oz, const uint8_t* valid, const double* oval,
                   double* Y, double* ybar, double* Yp, double* d);

/* Gaspari & Cohn (1999) eq. 4.10, z = dist / c, support 2c. */
double orc_gaspari_cohn(double z);

/*
 * LETKF analysis (LETKF.hpp:63-119, 152-243).  X is updated in place.
 * cols_sel: optional list of ncols_sel linear column ids (y*nx+x) to analyse (SNAPSHOT only;
 * other columns are left untouched); NULL = all columns.
 * counts_out: optional [nx*ny] local-obs count per analysed column (level-0 count when radius_v>0).
 * W_out: optional [ncols][k][k] row-major transform W (= wa 1^T + Wa; for REF_COMPAT the k-vector
 *        of per-member scale factors s_i is stored in row 0) of every analysed column, level 0.
 * Returns 0, or <0 on a numerical failure (non-SPD matrix).
 */
int orc_letkf(const orc_letkf_params* p, double* X, const int32_t* ox, const int32_t* oy,
              const int32_t* oz, const double* oval, const double* oerr, const uint8_t* valid,
              const int64_t* cols_sel, int64_t ncols_sel, int32_t* counts_out, double* W_out);

/* ---- GEOGRAPHIC locations and multi-variable states (the WRF-shaped case) ---------------------- */

/* Location::distance_to for two GEOGRAPHIC locations: haversine kilometres, R = 6371 km
 * (Location.hpp:213-217, 325, 333, 349-357). */
double orc_distance_cartesian(double x1, double y1, double z1, double x2, double y2, double z2);
double orc_distance_geo(double lat1, double lon1, double lat2, double lon2);

/* LETKF.hpp:159-165 with GEOGRAPHIC locations.  min_margin (optional, in/out): smallest |distance - radius|
 * seen -- the tests use it to show that no pair sits within rounding of the cutoff. */
int64_t orc_select_local_geo(double clat, double clon, int64_t P, const double* olat, const double* olon,
                             double radius, int32_t* idx_out, double* min_margin);
void orc_select_counts_geo(int nx, int ny, const double* glat, const double* glon, int64_t P,
                           const double* olat, const double* olon, double radius, int32_t* counts,
                           double* min_margin);

/* IdentityObsOperator::convertGeographicToGrid (IdentityObsOperator.hpp:484-530): nearest grid point (first
 * minimum of the Euclidean distance in degrees over glat/glon [ny][nx]) and nearest vertical level. */
void orc_geo_locate(int64_t P, const double* olat, const double* olon, const double* olev,
                    const double* glat, const double* glon, int nx, int ny, const double* vcoord,
                    int nlev, int32_t* ox, int32_t* oy, int32_t* oz);

typedef struct {
  const double *glat, *glon;   /* [ny][nx] column coordinates in degrees; NULL: GRID distances            */
  const double *olat, *olon;   /* [P] observation coordinates in degrees (with glat)                     */
  int nvar;                    /* variables of the state, member layout [var][lev][y][x]; 0: one variable */
  const int32_t* var_nlev;     /* [nvar] levels per variable (sum = nz)                                   */
  const int32_t* ovar;         /* [P] variable each observation observes, NULL: variable 0               */
  /* staggered grids: H is evaluated on this ensemble ([k][nz_obs][ny_obs][nx_obs], the grid that holds the     */
  /* observed variables -- var_nlev / ovar then describe it) instead of the analysed X; NULL: on X itself       */
  const double* Xobs;
  int nx_obs, ny_obs, nz_obs;
} orc_ext;

/* IdentityObsOperator::apply for one member of a multi-variable state at located observation coordinates
 * (IdentityObsOperator.hpp:154-180, 236-281, 681-711). */
void orc_hx_ext(const double* member, int nx, int ny, int nz, const orc_ext* ext, int64_t P, const int32_t* ox,
                const int32_t* oy, const int32_t* oz, const uint8_t* valid, double* out);

int orc_letkf_ext(const orc_letkf_params* p, const orc_ext* ext, double* X, const int32_t* ox,
                  const int32_t* oy, const int32_t* oz, const double* oval, const double* oerr,
                  const uint8_t* valid, const int64_t* cols_sel, int64_t ncols_sel, int32_t* counts_out,
                  double* W_out);

/* Global ETKF (ETKF.hpp:100-179). X in place. */
int orc_etkf(double* X, int nx, int ny, int nz, int k, int64_t P, const int32_t* ox,
             const int32_t* oy, const int32_t* oz, const double* oval, const double* oerr,
             const uint8_t* valid, double inflation);

typedef struct {
  double innovation_norm, background_spread, analysis_spread;
  double max_kalman_gain, min_kalman_gain, condition_number;
} orc_enkf_diag;

/* Global stochastic EnKF (EnKF.hpp:139-256) with supplied standard-normal draws Z (P x k,
 * row-major; obs_pert = sqrt(R_ii) * Z, EnKF.hpp:340-361).  X in place.
 * want_gain_stats: form K explicitly to get max/min (O(n*P) memory). */
int orc_lwenkf(double* X, int nx, int ny, int nz, int k, int64_t P, const int32_t* ox, const int32_t* oy,
               const int32_t* oz, const double* oval, const double* oerr, const uint8_t* valid, double inflation,
               double radius, int loc_fn, int weighting, const double* Z, double* diag);
int orc_enkf(double* X, int nx, int ny, int nz, int k, int64_t P, const int32_t* ox,
             const int32_t* oy, const int32_t* oz, const double* oval, const double* oerr,
             const uint8_t* valid, double inflation, const double* Z, int want_gain_stats,
             int nthreads, orc_enkf_diag* diag);

/* dense helpers exposed for unit tests (row-major k x k) */
int orc_lu_inverse(int k, const double* A, double* Ainv);
int orc_cholesky_lower(int k, const double* A, double* L);
int orc_jacobi_eigh(int k, const double* A, double* evals, double* V, int* sweeps);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
