/*
 * metada_oracle.c -- CPU restatement (plain C, FP64) of the METADA ensemble Kalman analysis path.
 *
 * TEST INFRASTRUCTURE ONLY -- see metada_oracle.h.  Not shipped, not on the product path.
 * Parity pin: checked against the reference's own headers built in oracle/_ref (tests/golden).
 *
 * Reference paths are relative to /root/reference/src.
 * Build: see oracle/Makefile (-O2 -ffp-contract=off so that H(x) keeps the operation-by-operation
 * rounding of the reference's plain -O2 build; OpenMP only across independent columns).
 */
#include "metada_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------ distance / selection */

/* framework/base/Location.hpp:204-211 (GRID x GRID: 2-D Euclid on integer indices, level ignored) */
double orc_distance_grid(int i1, int j1, int i2, int j2) {
  double dx = (double)(i1 - i2);
  double dy = (double)(j1 - j2);
  return sqrt(dx * dx + dy * dy);
}

/* framework/algorithms/LETKF.hpp:159-165 : ascending obs index, inclusive <= */
int64_t orc_select_local(int gx, int gy, int64_t P, const int32_t* ox, const int32_t* oy,
                         double radius, int32_t* idx_out) {
  int64_t c = 0;
  for (int64_t i = 0; i < P; ++i) {
    double distance = orc_distance_grid(gx, gy, ox[i], oy[i]);
    if (distance <= radius) {
      if (idx_out) idx_out[c] = (int32_t)i;
      ++c;
    }
  }
  return c;
}

void orc_select_counts(int nx, int ny, int64_t P, const int32_t* ox, const int32_t* oy,
                       double radius, int32_t* counts, int nthreads) {
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
#endif
  for (int64_t g = 0; g < (int64_t)nx * ny; ++g) {
    int gx = (int)(g % nx), gy = (int)(g / nx);
    counts[g] = (int32_t)orc_select_local(gx, gy, P, ox, oy, radius, NULL);
  }
  (void)nthreads;
}

/* framework/base/Location.hpp:213-217 (GEOGRAPHIC x GEOGRAPHIC: horizontal great circle, level ignored) with
 * deg2rad (:333) and haversine (:349-357), kEarthRadiusKm = 6371.0 (:325); result in kilometres */
static double orc_deg2rad(double deg) { return deg * 3.14159265358979323846 / 180.0; }
double orc_distance_geo(double lat1, double lon1, double lat2, double lon2) {
  double dlat = orc_deg2rad(lat2 - lat1);
  double dlon = orc_deg2rad(lon2 - lon1);
  double a = sin(dlat / 2) * sin(dlat / 2) +
             cos(orc_deg2rad(lat1)) * cos(orc_deg2rad(lat2)) * sin(dlon / 2) * sin(dlon / 2);
  double c = 2 * atan2(sqrt(a), sqrt(1 - a));
  return 6371.0 * c;
}

/* Location::distance_to, CARTESIAN (Location.hpp:217-225): 3-D Euclid, sum in the order dx^2 + dy^2 + dz^2.
 * Restated and pinned for completeness of distance_to; no filter can reach it in the reference -- H throws for a
 * CARTESIAN observation (IdentityObsOperator.hpp:251-255 -> Location.hpp:100-103) -- so there is no device path. */
double orc_distance_cartesian(double x1, double y1, double z1, double x2, double y2, double z2) {
  double dx = x1 - x2, dy = y1 - y2, dz = z1 - z2;
  return sqrt(dx * dx + dy * dy + dz * dz);
}

/* LETKF.hpp:159-165 with GEOGRAPHIC locations: ascending obs index, inclusive <= (kilometres) */
int64_t orc_select_local_geo(double clat, double clon, int64_t P, const double* olat, const double* olon,
                             double radius, int32_t* idx_out, double* min_margin) {
  int64_t c = 0;
  for (int64_t i = 0; i < P; ++i) {
    double distance = orc_distance_geo(clat, clon, olat[i], olon[i]);
    if (min_margin) {
      double m = fabs(distance - radius);
      if (m < *min_margin) *min_margin = m;
    }
    if (distance <= radius) {
      if (idx_out) idx_out[c] = (int32_t)i;
      ++c;
    }
  }
  return c;
}

void orc_select_counts_geo(int nx, int ny, const double* glat, const double* glon, int64_t P,
                           const double* olat, const double* olon, double radius, int32_t* counts,
                           double* min_margin) {
  double mm = INFINITY;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) reduction(min : mm)
#endif
  for (int64_t g = 0; g < (int64_t)nx * ny; ++g) {
    double m = INFINITY;
    counts[g] = (int32_t)orc_select_local_geo(glat[g], glon[g], P, olat, olon, radius, NULL, &m);
    if (m < mm) mm = m;
  }
  if (min_margin) *min_margin = mm;
}

/* backends/common/obsoperator/IdentityObsOperator.hpp:484-530 (convertGeographicToGrid, the branch for a
 * geometry with 2-D coordinate arrays): FIRST minimum of the Euclidean distance in degrees over the grid in
 * linear (y-outer, x-inner) order, then the first minimum of |level - vertical_coords[z]| (k = 0 without
 * vertical coordinates). */
void orc_geo_locate(int64_t P, const double* olat, const double* olon, const double* olev,
                    const double* glat, const double* glon, int nx, int ny, const double* vcoord,
                    int nlev, int32_t* ox, int32_t* oy, int32_t* oz) {
  const int64_t G = (int64_t)nx * ny;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int64_t i = 0; i < P; ++i) {
    const double lat = olat[i], lon = olon[i];
    double min_dist = DBL_MAX;
    int64_t min_idx = 0;
    for (int64_t idx = 0; idx < G; ++idx) {
      double grid_lon = glon[idx], grid_lat = glat[idx];
      double dist = sqrt((lon - grid_lon) * (lon - grid_lon) + (lat - grid_lat) * (lat - grid_lat));
      if (dist < min_dist) { min_dist = dist; min_idx = idx; }
    }
    int k = 0;
    if (vcoord && nlev > 0) {
      double level = olev ? olev[i] : 0.0, min_vert_dist = DBL_MAX;
      for (int z = 0; z < nlev; ++z) {
        double dist = fabs(level - vcoord[z]);
        if (dist < min_vert_dist) { min_vert_dist = dist; k = z; }
      }
    }
    ox[i] = (int32_t)(min_idx % nx); oy[i] = (int32_t)(min_idx / nx); oz[i] = k;
  }
}

/* ------------------------------------------------------------------ H(x): 4-point IDW */

/* backends/common/obsoperator/IdentityObsOperator.hpp:594-638 (find4NearestGridPoints) and
 * :643-676 (idw4Interpolation), linear index kk*(ny*nx) + jj*nx + ii. */
/* Multi-variable states (:681-711, idw4InterpolationVariable): the neighbours are searched on the geometry's
 * nz levels, the value is read from the observation's variable -- `s` points at that variable's block and
 * var_nlev is its level count (a 2-D variable ignores the level, :706-707). */
static double hx_one_var(const double* s, int nx, int ny, int nz, int var_nlev, int oxi, int oyi, int ozi) {
  double x = (double)oxi, y = (double)oyi, z = (double)ozi;
  x = fmax(0.0, fmin((double)(nx - 1), x));
  y = fmax(0.0, fmin((double)(ny - 1), y));
  z = fmax(0.0, fmin((double)(nz - 1), z));
  size_t i0 = (size_t)floor(x), j0 = (size_t)floor(y), k0 = (size_t)floor(z);
  size_t i1 = i0 + 1 < (size_t)(nx - 1) ? i0 + 1 : (size_t)(nx - 1);
  size_t j1 = j0 + 1 < (size_t)(ny - 1) ? j0 + 1 : (size_t)(ny - 1);
  size_t k1 = k0 + 1 < (size_t)(nz - 1) ? k0 + 1 : (size_t)(nz - 1);
  size_t ii[8], jj[8], kk[8];
  double dist[8];
  int cnt;
  if (nz == 1) {
    size_t ci[4] = {i0, i1, i0, i1}, cj[4] = {j0, j0, j1, j1};
    for (int c = 0; c < 4; ++c) {
      ii[c] = ci[c]; jj[c] = cj[c]; kk[c] = 0;
      dist[c] = sqrt((x - ii[c]) * (x - ii[c]) + (y - jj[c]) * (y - jj[c]));
    }
    cnt = 4;
  } else {
    size_t ci[8] = {i0, i1, i0, i1, i0, i1, i0, i1};
    size_t cj[8] = {j0, j0, j1, j1, j0, j0, j1, j1};
    size_t ck[8] = {k0, k0, k0, k0, k1, k1, k1, k1};
    for (int c = 0; c < 8; ++c) {
      ii[c] = ci[c]; jj[c] = cj[c]; kk[c] = ck[c];
      dist[c] = sqrt((x - ii[c]) * (x - ii[c]) + (y - jj[c]) * (y - jj[c]) +
                     (z - kk[c]) * (z - kk[c]));
    }
    /* std::sort on 8 elements with strict '<' == libstdc++ insertion sort == stable */
    for (int a = 1; a < 8; ++a) {
      double dv = dist[a]; size_t iv = ii[a], jv = jj[a], kv = kk[a];
      int b = a - 1;
      while (b >= 0 && dv < dist[b]) {
        dist[b + 1] = dist[b]; ii[b + 1] = ii[b]; jj[b + 1] = jj[b]; kk[b + 1] = kk[b];
        --b;
      }
      dist[b + 1] = dv; ii[b + 1] = iv; jj[b + 1] = jv; kk[b + 1] = kv;
    }
    cnt = 4;
  }
  double weighted_sum = 0.0, weight_sum = 0.0;
  for (int c = 0; c < cnt; ++c) {
    double w = (dist[c] == 0.0) ? 1e12 : 1.0 / dist[c];
    size_t linear_index = (var_nlev > 1 ? kk[c] : 0) * ((size_t)ny * nx) + jj[c] * (size_t)nx + ii[c];
    weighted_sum += w * s[linear_index];
    weight_sum += w;
  }
  return weighted_sum / weight_sum;
}

static double hx_one(const double* s, int nx, int ny, int nz, int oxi, int oyi, int ozi) {
  return hx_one_var(s, nx, ny, nz, nz, oxi, oyi, ozi);
}

/* IdentityObsOperator.hpp:154-180 : invalid obs -> 0.0 */
void orc_hx_idw4(const double* member, int nx, int ny, int nz, int64_t P, const int32_t* ox,
                 const int32_t* oy, const int32_t* oz, const uint8_t* valid, double* out) {
  for (int64_t i = 0; i < P; ++i) {
    if (valid && !valid[i]) { out[i] = 0.0; continue; }
    out[i] = hx_one(member, nx, ny, nz, ox[i], oy[i], oz ? oz[i] : 0);
  }
}

/* framework/adapters/Ensemble.hpp:105-114 */
void orc_ensemble_mean(const double* X, int k, int64_t n, double* mean) {
  for (int64_t i = 0; i < n; ++i) mean[i] = 0.0;
  for (int m = 0; m < k; ++m)
    for (int64_t i = 0; i < n; ++i) mean[i] += X[(int64_t)m * n + i];
  double f = 1.0 / (double)k;
  for (int64_t i = 0; i < n; ++i) mean[i] *= f;
}

/* LETKF.hpp:197-211 / ETKF.hpp:128-141 / EnKF.hpp:168-181 */
void orc_obs_space(const double* X, int nx, int ny, int nz, int k, int64_t P, const int32_t* ox,
                   const int32_t* oy, const int32_t* oz, const uint8_t* valid, const double* oval,
                   double* Y, double* ybar, double* Yp, double* d) {
  int64_t n = (int64_t)nx * ny * nz;
  double* col = (double*)malloc(sizeof(double) * (size_t)(P > 0 ? P : 1));
  for (int m = 0; m < k; ++m) {
    orc_hx_idw4(X + (int64_t)m * n, nx, ny, nz, P, ox, oy, oz, valid, col);
    for (int64_t i = 0; i < P; ++i) Y[i * k + m] = col[i];
  }
  free(col);
  for (int64_t i = 0; i < P; ++i) {
    double s = 0.0;
    for (int m = 0; m < k; ++m) s += Y[i * k + m];
    double mean = s / (double)k;               /* Eigen rowwise().mean() = sum / size */
    if (ybar) ybar[i] = mean;
    if (Yp) for (int m = 0; m < k; ++m) Yp[i * k + m] = Y[i * k + m] - mean;
    if (d) d[i] = oval[i] - mean;
  }
}

/* Y, Y', d for a multi-variable state: member layout [var][lev][y][x], observation i reads variable ovar[i]
 * (IdentityObsOperator.hpp:236-281, 681-711); nzg = levels of the geometry (the smallest multi-level variable; a vertically
 * staggered one has nzg + 1). */
static void obs_space_ext(const double* X, int nx, int ny, int nz, int k, int64_t P, const int32_t* ox,
                          const int32_t* oy, const int32_t* oz, const uint8_t* valid, const double* oval,
                          const orc_ext* ext, double* Y, double* Yp, double* d) {
  const int64_t G = (int64_t)nx * ny, n = G * nz;
  /* geometry levels = the mass-level variables' count = the smallest multi-level one; a vertically staggered variable
   * (W: nz + 1 levels) is read with its own dimensions but searched on the geometry's levels (:684-711) */
  int off[64] = {0}, nzg = 0;
  for (int v = 0; v < ext->nvar; ++v) {
    off[v + 1] = off[v] + ext->var_nlev[v];
    if (ext->var_nlev[v] > 1 && (nzg == 0 || ext->var_nlev[v] < nzg)) nzg = ext->var_nlev[v];
  }
  if (nzg == 0) nzg = 1;
  for (int m = 0; m < k; ++m)
    for (int64_t i = 0; i < P; ++i) {
      if (valid && !valid[i]) { Y[i * k + m] = 0.0; continue; }
      const int v = ext->ovar ? ext->ovar[i] : 0;
      Y[i * k + m] = hx_one_var(X + (int64_t)m * n + (int64_t)off[v] * G, nx, ny, nzg, ext->var_nlev[v], ox[i],
                                oy[i], oz ? oz[i] : 0);
    }
  for (int64_t i = 0; i < P; ++i) {
    double s = 0.0;
    for (int m = 0; m < k; ++m) s += Y[i * k + m];
    double mean = s / (double)k;
    for (int m = 0; m < k; ++m) Yp[i * k + m] = Y[i * k + m] - mean;
    d[i] = oval[i] - mean;
  }
}

/* H(x) of ONE member of a multi-variable state at located (integer) observation coordinates: the body of
 * obs_space_ext exposed for the pin against the reference's IdentityObsOperator (oracle/ref_obsop_geo.cpp). */
void orc_hx_ext(const double* member, int nx, int ny, int nz, const orc_ext* ext, int64_t P, const int32_t* ox,
                const int32_t* oy, const int32_t* oz, const uint8_t* valid, double* out) {
  const int64_t G = (int64_t)nx * ny;
  int off[64] = {0}, nzg = 0;
  (void)nz;
  for (int v = 0; v < ext->nvar && v < 63; ++v) {
    off[v + 1] = off[v] + ext->var_nlev[v];
    if (ext->var_nlev[v] > 1 && (nzg == 0 || ext->var_nlev[v] < nzg)) nzg = ext->var_nlev[v];
  }
  if (nzg == 0) nzg = 1;
  for (int64_t i = 0; i < P; ++i) {
    if (valid && !valid[i]) { out[i] = 0.0; continue; }
    const int v = ext->ovar ? ext->ovar[i] : 0;
    out[i] = hx_one_var(member + (int64_t)off[v] * G, nx, ny, nzg, ext->var_nlev[v], ox[i], oy[i], oz ? oz[i] : 0);
  }
}

double orc_gaspari_cohn(double z) {
  z = fabs(z);
  if (z >= 2.0) return 0.0;
  if (z <= 1.0)
    return (((-0.25 * z + 0.5) * z + 0.625) * z - 5.0 / 3.0) * z * z + 1.0;
  /* close to z = 2 the polynomial cancels to a few 1e-16 of either sign; the taper itself is >= 0 */
  return fmax(((((z / 12.0 - 0.5) * z + 0.625) * z + 5.0 / 3.0) * z - 5.0) * z + 4.0 - 2.0 / (3.0 * z), 0.0);
}

/* LWEnKF.hpp:597-621 (switch over LocalizationFunction) and :624-635 (the reference's own
 * "Gaspari-Cohn" polynomial, which is not the Gaspari-Cohn taper: 4 at r = 0, jump at r = 1). */
double orc_loc_weight(int loc, double dist, double support, double scale) {
  switch (loc) {
    case ORC_LOC_GASPARI_COHN: return orc_gaspari_cohn(dist / (0.5 * support));
    case ORC_LOC_GAUSSIAN: { double r = dist / scale; return exp(-0.5 * r * r); }   /* :601-602 */
    case ORC_LOC_EXPONENTIAL: return exp(-(dist / scale));                           /* :604-605 */
    case ORC_LOC_REF_GASPARI_COHN: {                                                 /* :624-635 */
      double r = dist / scale;
      if (r >= 2.0) return 0.0;
      if (r >= 1.0) { double z = r - 1.0; return ((-0.25 * z + 0.5) * z + 0.625) * z + 0.125; }
      return (((-0.25 * r + 0.5) * r + 0.625) * r - 5.0) * r + 4.0;
    }
    default: return 1.0;
  }
}

/* Metrics.hpp, loop for loop: mean :108-121, spread :137-150, bias :166-173, correlation :189-210,
 * CRPS :232-254, RMSE :265-273, average spread :286-292. */
void orc_metrics(const double* X, const double* truth, int64_t n, int k, double* mean_out,
                 double* spread_out, double out[5]) {
  double* mean = (double*)calloc((size_t)n, sizeof(double));
  double* spread = (double*)calloc((size_t)n, sizeof(double));
  for (int m = 0; m < k; ++m)
    for (int64_t i = 0; i < n; ++i) mean[i] += X[(int64_t)m * n + i];
  for (int64_t i = 0; i < n; ++i) mean[i] /= (double)k;      /* val /= ens_size (size_t -> double) */
  for (int m = 0; m < k; ++m)
    for (int64_t i = 0; i < n; ++i) {
      double diff = X[(int64_t)m * n + i] - mean[i];
      spread[i] += diff * diff;
    }
  for (int64_t i = 0; i < n; ++i) spread[i] = sqrt(spread[i] / (double)(k - 1));
  double rmse = 0.0, bias = 0.0;
  for (int64_t i = 0; i < n; ++i) { double diff = mean[i] - truth[i]; rmse += diff * diff; }
  for (int64_t i = 0; i < n; ++i) bias += mean[i] - truth[i];
  double sx = 0.0, sy = 0.0, sxy = 0.0, sx2 = 0.0, sy2 = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    sx += mean[i]; sy += truth[i]; sxy += mean[i] * truth[i];
    sx2 += mean[i] * mean[i]; sy2 += truth[i] * truth[i];
  }
  double nn = (double)n;
  double crps = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    double sum_diff = 0.0, sum_truth_diff = 0.0;
    for (int j = 0; j < k; ++j)
      for (int l = 0; l < k; ++l) sum_diff += fabs(X[(int64_t)j * n + i] - X[(int64_t)l * n + i]);
    for (int j = 0; j < k; ++j) sum_truth_diff += fabs(X[(int64_t)j * n + i] - truth[i]);
    crps += sum_truth_diff / (double)k - sum_diff / (2.0 * (double)k * (double)k);
  }
  double avg = 0.0;
  for (int64_t i = 0; i < n; ++i) avg += spread[i];
  out[0] = sqrt(rmse / nn);
  out[1] = bias / nn;
  out[2] = (nn * sxy - sx * sy) / sqrt((nn * sx2 - sx * sx) * (nn * sy2 - sy * sy));
  out[3] = crps / nn;
  out[4] = avg / nn;
  if (mean_out) memcpy(mean_out, mean, sizeof(double) * (size_t)n);
  if (spread_out) memcpy(spread_out, spread, sizeof(double) * (size_t)n);
  free(mean); free(spread);
}

/* ------------------------------------------------------------------ dense kit (row-major) */

/* Partial-pivot LU inverse: what Eigen's MatrixXd::inverse() does for dynamic sizes. */
int orc_lu_inverse(int k, const double* A, double* Ainv) {
  double* M = (double*)malloc(sizeof(double) * (size_t)k * k);
  int* piv = (int*)malloc(sizeof(int) * (size_t)k);
  memcpy(M, A, sizeof(double) * (size_t)k * k);
  for (int i = 0; i < k; ++i) piv[i] = i;
  for (int c = 0; c < k; ++c) {
    int pr = c; double best = fabs(M[c * k + c]);
    for (int r = c + 1; r < k; ++r) if (fabs(M[r * k + c]) > best) { best = fabs(M[r * k + c]); pr = r; }
    if (best == 0.0) { free(M); free(piv); return -1; }
    if (pr != c) {
      for (int j = 0; j < k; ++j) { double t = M[c * k + j]; M[c * k + j] = M[pr * k + j]; M[pr * k + j] = t; }
      int t = piv[c]; piv[c] = piv[pr]; piv[pr] = t;
    }
    for (int r = c + 1; r < k; ++r) {
      double f = M[r * k + c] / M[c * k + c];
      M[r * k + c] = f;
      for (int j = c + 1; j < k; ++j) M[r * k + j] -= f * M[c * k + j];
    }
  }
  /* solve L U X = P I, column by column */
  double* y = (double*)malloc(sizeof(double) * (size_t)k);
  for (int col = 0; col < k; ++col) {
    for (int i = 0; i < k; ++i) {
      double s = (piv[i] == col) ? 1.0 : 0.0;
      for (int j = 0; j < i; ++j) s -= M[i * k + j] * y[j];
      y[i] = s;
    }
    for (int i = k - 1; i >= 0; --i) {
      double s = y[i];
      for (int j = i + 1; j < k; ++j) s -= M[i * k + j] * Ainv[j * k + col];
      Ainv[i * k + col] = s / M[i * k + i];
    }
  }
  free(y); free(M); free(piv);
  return 0;
}

/* Standard lower Cholesky (Eigen llt().matrixL()); upper triangle of L is zero. */
int orc_cholesky_lower(int k, const double* A, double* L) {
  memset(L, 0, sizeof(double) * (size_t)k * k);
  for (int j = 0; j < k; ++j) {
    double s = A[j * k + j];
    for (int t = 0; t < j; ++t) s -= L[j * k + t] * L[j * k + t];
    if (!(s > 0.0)) return -1;
    double ljj = sqrt(s);
    L[j * k + j] = ljj;
    for (int i = j + 1; i < k; ++i) {
      double v = A[i * k + j];
      for (int t = 0; t < j; ++t) v -= L[i * k + t] * L[j * k + t];
      L[i * k + j] = v / ljj;
    }
  }
  return 0;
}

/* Cyclic two-sided Jacobi; eigenvalues unsorted; V columns = eigenvectors. */
int orc_jacobi_eigh(int k, const double* Ain, double* evals, double* V, int* sweeps_out) {
  double* A = (double*)malloc(sizeof(double) * (size_t)k * k);
  memcpy(A, Ain, sizeof(double) * (size_t)k * k);
  for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) V[i * k + j] = (i == j) ? 1.0 : 0.0;
  int sweep = 0;
  for (; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < k; ++i) {
      diag += A[i * k + i] * A[i * k + i];
      for (int j = i + 1; j < k; ++j) off += A[i * k + j] * A[i * k + j];
    }
    if (off <= 1e-30 * diag || off == 0.0) break;
    for (int p = 0; p < k - 1; ++p)
      for (int q = p + 1; q < k; ++q) {
        double apq = A[p * k + q];
        if (apq == 0.0) continue;
        double app = A[p * k + p], aqq = A[q * k + q];
        if (fabs(apq) < 1e-300) continue;
        double theta = (aqq - app) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int r = 0; r < k; ++r) {
          double arp = A[r * k + p], arq = A[r * k + q];
          A[r * k + p] = c * arp - s * arq;
          A[r * k + q] = s * arp + c * arq;
        }
        for (int r = 0; r < k; ++r) {
          double apr = A[p * k + r], aqr = A[q * k + r];
          A[p * k + r] = c * apr - s * aqr;
          A[q * k + r] = s * apr + c * aqr;
        }
        for (int r = 0; r < k; ++r) {
          double vrp = V[r * k + p], vrq = V[r * k + q];
          V[r * k + p] = c * vrp - s * vrq;
          V[r * k + q] = s * vrp + c * vrq;
        }
      }
  }
  for (int i = 0; i < k; ++i) evals[i] = A[i * k + i];
  if (sweeps_out) *sweeps_out = sweep;
  free(A);
  return sweep < 60 ? 0 : -2;
}

/* ------------------------------------------------------------------ local transform */

typedef struct {
  double *A, *Pa, *L, *T, *V, *ev, *g, *wa, *Wa;
} xform_ws;

static void ws_alloc(xform_ws* w, int k, int64_t pmax) {
  size_t kk = (size_t)k * k;
  w->A = (double*)malloc(sizeof(double) * kk);
  w->Pa = (double*)malloc(sizeof(double) * kk);
  w->L = (double*)malloc(sizeof(double) * kk);
  w->V = (double*)malloc(sizeof(double) * kk);
  w->Wa = (double*)malloc(sizeof(double) * kk);
  w->T = (double*)malloc(sizeof(double) * (size_t)k * (size_t)(pmax > 0 ? pmax : 1));
  w->ev = (double*)malloc(sizeof(double) * (size_t)k);
  w->g = (double*)malloc(sizeof(double) * (size_t)k);
  w->wa = (double*)malloc(sizeof(double) * (size_t)k);
}
static void ws_free(xform_ws* w) {
  free(w->A); free(w->Pa); free(w->L); free(w->V); free(w->Wa); free(w->T); free(w->ev);
  free(w->g); free(w->wa);
}

/*
 * Local ensemble transform from the local obs rows.
 *   Yl  [pl][k] perturbations, dl [pl] innovations, rinv [pl] (rho_i / sigma_i^2; REF_COMPAT: unused)
 * Output: w->wa [k], w->Wa [k][k] (row-major).
 *   REF_COMPAT (LETKF.hpp:214-224): A = Y'^T I Y' + (k-1) I ; Pa = A^-1 * infl ; wa = ((Pa Y'^T) I) d ;
 *                                   Wa = sqrt(k-1) chol(Pa)
 *   REF_ETKF   (ETKF.hpp:150-159):  A = Y'^T R^-1 Y' + (k-1) I ; Pa = A^-1 ; wa = Pa Y'^T R^-1 d ;
 *                                   Wa = sqrt(k-1) chol(Pa)
 *   CANONICAL  (Hunt et al. 2007, eqs 21-24): A = (k-1)/infl I + Y'^T (rho R^-1) Y' = V L V^T ;
 *                                   wa = V L^-1 V^T Y'^T (rho R^-1) d ; Wa = V sqrt((k-1)/L) V^T
 */
static int local_transform(int mode, double infl, int k, int64_t pl, const double* Yl,
                           const double* dl, const double* rinv, xform_ws* w) {
  double *A = w->A, *Pa = w->Pa;
  const double km1 = (double)(k - 1);
  for (int a = 0; a < k; ++a)
    for (int b = 0; b < k; ++b) {
      double s = 0.0;
      if (mode == ORC_MODE_REF_COMPAT)
        for (int64_t i = 0; i < pl; ++i) s += Yl[i * k + a] * Yl[i * k + b];
      else
        for (int64_t i = 0; i < pl; ++i) s += (Yl[i * k + a] * rinv[i]) * Yl[i * k + b];
      A[a * k + b] = s;
    }
  if (mode == ORC_MODE_CANONICAL) {
    for (int a = 0; a < k; ++a) A[a * k + a] += km1 / infl;
    for (int a = 0; a < k; ++a) {
      double s = 0.0;
      for (int64_t i = 0; i < pl; ++i) s += Yl[i * k + a] * (rinv[i] * dl[i]);
      w->g[a] = s;
    }
    int sweeps;
    if (orc_jacobi_eigh(k, A, w->ev, w->V, &sweeps)) return -2;
    const double* V = w->V;
    /* t = V^T g ; wa = V (t / ev) */
    for (int c = 0; c < k; ++c) {
      double s = 0.0;
      for (int a = 0; a < k; ++a) s += V[a * k + c] * w->g[a];
      w->L[c] = s / w->ev[c];
    }
    for (int a = 0; a < k; ++a) {
      double s = 0.0;
      for (int c = 0; c < k; ++c) s += V[a * k + c] * w->L[c];
      w->wa[a] = s;
    }
    for (int a = 0; a < k; ++a)
      for (int b = 0; b < k; ++b) {
        double s = 0.0;
        for (int c = 0; c < k; ++c) s += V[a * k + c] * sqrt(km1 / w->ev[c]) * V[b * k + c];
        w->Wa[a * k + b] = s;
      }
    return 0;
  }
  for (int a = 0; a < k; ++a) A[a * k + a] += km1;
  if (orc_lu_inverse(k, A, Pa)) return -1;
  if (mode == ORC_MODE_REF_COMPAT)
    for (int a = 0; a < k * k; ++a) Pa[a] *= infl;
  /* T = Pa Y'^T (R^-1) : k x pl ; wa = T d */
  for (int a = 0; a < k; ++a) {
    for (int64_t i = 0; i < pl; ++i) {
      double s = 0.0;
      for (int b = 0; b < k; ++b) s += Pa[a * k + b] * Yl[i * k + b];
      w->T[a * pl + i] = (mode == ORC_MODE_REF_COMPAT) ? s : s * rinv[i];
    }
    double s = 0.0;
    for (int64_t i = 0; i < pl; ++i) s += w->T[a * pl + i] * dl[i];
    w->wa[a] = s;
  }
  if (orc_cholesky_lower(k, Pa, w->L)) return -1;
  const double sq = sqrt(km1);
  for (int a = 0; a < k * k; ++a) w->Wa[a] = sq * w->L[a];
  return 0;
}

/* Apply the transform to the k values of one grid point.
 *   REF_COMPAT (LETKF.hpp:227-238): xa_i = (m + xp_i wa_i) + sum_j xp_i Wa_ij   (asDiagonal form)
 *   REF_ETKF   (ETKF.hpp:125,163-169): xp *= infl ; xa_i = (m + sum_j xp_j wa_j) + sum_j xp_j Wa_ji
 *   CANONICAL: as REF_ETKF without the pre-scaling (inflation is inside A). */
static void apply_point(int mode, double infl, int k, const xform_ws* w, double* x /*k, in/out*/) {
  double m = 0.0;
  for (int i = 0; i < k; ++i) m += x[i];
  m /= (double)k; /* Eigen .mean() = sum / size */
  double xp[1024];
  for (int i = 0; i < k; ++i) xp[i] = x[i] - m;
  if (mode == ORC_MODE_REF_COMPAT) {
    for (int i = 0; i < k; ++i) {
      double xam = m + xp[i] * w->wa[i];
      double rs = 0.0;
      for (int j = 0; j < k; ++j) rs += xp[i] * w->Wa[i * k + j];
      x[i] = xam + rs;
    }
    return;
  }
  if (mode == ORC_MODE_REF_ETKF)
    for (int i = 0; i < k; ++i) xp[i] *= infl;
  double inc = 0.0;
  for (int j = 0; j < k; ++j) inc += xp[j] * w->wa[j];
  double xam = m + inc;
  for (int i = 0; i < k; ++i) {
    double s = 0.0;
    for (int j = 0; j < k; ++j) s += xp[j] * w->Wa[j * k + i];
    x[i] = xam + s;
  }
}

/* LETKF.hpp:167-190 : no local obs -> inflate perturbations by sqrt(inflation)
 * (REF_ETKF has no such branch in ETKF.hpp; with p=0 its formulas give Wa = I, xp *= infl.) */
static void inflate_point(int mode, double infl, int k, double* x) {
  double m = 0.0;
  for (int i = 0; i < k; ++i) m += x[i];
  m /= (double)k;
  double f = (mode == ORC_MODE_REF_ETKF) ? infl : sqrt(infl);
  for (int i = 0; i < k; ++i) {
    double xp = x[i] - m;
    xp *= f;
    x[i] = m + xp;
  }
}

static void store_W(int mode, int k, const xform_ws* w, double* Wo) {
  if (mode == ORC_MODE_REF_COMPAT) {
    memset(Wo, 0, sizeof(double) * (size_t)k * k);
    for (int i = 0; i < k; ++i) {
      double rs = 0.0;
      for (int j = 0; j < k; ++j) rs += w->Wa[i * k + j];
      Wo[i] = w->wa[i] + rs;
    }
  } else {
    for (int j = 0; j < k; ++j)
      for (int i = 0; i < k; ++i) Wo[j * k + i] = w->wa[j] + w->Wa[j * k + i];
  }
}

/* ------------------------------------------------------------------ LETKF drivers */

static int letkf_snapshot(const orc_letkf_params* p, const orc_ext* ext, double* X, const int32_t* ox,
                          const int32_t* oy, const int32_t* oz, const double* oval,
                          const double* oerr, const uint8_t* valid, const int64_t* cols_sel,
                          int64_t ncols_sel, int32_t* counts_out, double* W_out) {
  const int nx = p->nx, ny = p->ny, nz = p->nz, k = p->k;
  const int64_t P = p->P, G = (int64_t)nx * ny, n = G * nz;
  if (k > 1024) return -3;
  double* Y = (double*)malloc(sizeof(double) * (size_t)(P > 0 ? P : 1) * k);
  double* Yp = (double*)malloc(sizeof(double) * (size_t)(P > 0 ? P : 1) * k);
  double* d = (double*)malloc(sizeof(double) * (size_t)(P > 0 ? P : 1));
  const int geo = ext && ext->glat;            /* GEOGRAPHIC locations: haversine kilometres */
  int* levmap = (int*)malloc(sizeof(int) * (size_t)nz);   /* level of a state level inside its variable */
  for (int l = 0; l < nz; ++l) levmap[l] = l;
  if (ext && ext->Xobs) {
    /* Staggered grids: the analysed ensemble X lives on its own column set (e.g. the U grid, with its own glat /
     * glon), the observations are operated on the ensemble Xobs of the grid that holds the observed variables
     * (nx_obs x ny_obs, nz_obs levels in all; var_nlev / ovar describe ITS variables).  X is one variable. */
    if (ext->nvar > 0) {
      obs_space_ext(ext->Xobs, ext->nx_obs, ext->ny_obs, ext->nz_obs, k, P, ox, oy, oz, valid, oval, ext, Y, Yp, d);
    } else {
      orc_obs_space(ext->Xobs, ext->nx_obs, ext->ny_obs, ext->nz_obs, k, P, ox, oy, oz, valid, oval, Y, NULL, Yp, d);
    }
  } else if (ext && ext->nvar > 0) {
    int tot = 0;
    for (int v = 0; v < ext->nvar; ++v)
      for (int l = 0; l < ext->var_nlev[v]; ++l) if (tot < nz) levmap[tot++] = l;
    if (tot != nz) { free(Y); free(Yp); free(d); free(levmap); return -4; }
    obs_space_ext(X, nx, ny, nz, k, P, ox, oy, oz, valid, oval, ext, Y, Yp, d);
  } else {
    orc_obs_space(X, nx, ny, nz, k, P, ox, oy, oz, valid, oval, Y, NULL, Yp, d);
  }
  free(Y);
  const int64_t ncols = cols_sel ? ncols_sel : G;
  int rc_all = 0;
  int nthreads = p->nthreads;
#ifdef _OPENMP
  const double loc_scale = p->loc_scale > 0.0 ? p->loc_scale : p->radius;
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
  {
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(P > 0 ? P : 1));
    int32_t* idl = (int32_t*)malloc(sizeof(int32_t) * (size_t)(P > 0 ? P : 1));
    double* Yl = NULL; double* dl = NULL; double* rinv = NULL; int64_t cap = 0;
    xform_ws w; int64_t wcap = 64; ws_alloc(&w, k, wcap);
    double xk[1024];
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
    for (int64_t ci = 0; ci < ncols; ++ci) {
      const int64_t g = cols_sel ? cols_sel[ci] : ci;
      const int gx = (int)(g % nx), gy = (int)(g / nx);
      const int64_t ph = geo ? orc_select_local_geo(ext->glat[g], ext->glon[g], P, ext->olat, ext->olon,
                                                    p->radius, idx, NULL)
                             : orc_select_local(gx, gy, P, ox, oy, p->radius, idx);
      if (counts_out) counts_out[g] = (int32_t)ph;
      if (ph > cap) {
        cap = ph * 2; free(Yl); free(dl); free(rinv);
        Yl = (double*)malloc(sizeof(double) * (size_t)cap * k);
        dl = (double*)malloc(sizeof(double) * (size_t)cap);
        rinv = (double*)malloc(sizeof(double) * (size_t)cap);
      }
      if (ph > wcap) { ws_free(&w); wcap = ph * 2; ws_alloc(&w, k, wcap); }
      const int per_level = (p->radius_v > 0.0);
      const int nxf = per_level ? nz : 1;
      for (int lt = 0; lt < nxf; ++lt) {
        /* build the local set for this transform */
        int64_t pl = 0;
        for (int64_t a = 0; a < ph; ++a) {
          const int32_t i = idx[a];
          double rho = 1.0;
          if (per_level) {
            double dv = fabs((double)(oz[i] - levmap[lt]));
            if (!(dv <= p->radius_v)) continue;
            if (p->mode == ORC_MODE_CANONICAL && p->loc != ORC_LOC_CUTOFF)
              rho *= orc_loc_weight(p->loc, dv, p->radius_v, p->radius_v * (loc_scale / p->radius));
          }
          if (p->mode == ORC_MODE_CANONICAL && p->loc != ORC_LOC_CUTOFF)
            rho *= orc_loc_weight(p->loc,
                                  geo ? orc_distance_geo(ext->glat[g], ext->glon[g], ext->olat[i], ext->olon[i])
                                      : orc_distance_grid(gx, gy, ox[i], oy[i]),
                                  p->radius, loc_scale);
          idl[pl] = i;
          double var = oerr[i] * oerr[i]; /* GridObservation.hpp:239-252 */
          if (valid && !valid[i]) var = INFINITY;
          if (p->mode == ORC_MODE_CANONICAL && !p->use_R) var = 1.0;
          rinv[pl] = rho / var;
          dl[pl] = d[i];
          memcpy(Yl + pl * k, Yp + (int64_t)i * k, sizeof(double) * (size_t)k);
          ++pl;
        }
        int have = 0;
        if (pl > 0) {
          int rc = local_transform(p->mode, p->inflation, k, pl, Yl, dl, rinv, &w);
          if (rc) {
#ifdef _OPENMP
#pragma omp critical
#endif
            rc_all = rc;
            continue;
          }
          have = 1;
          if (W_out && lt == 0) store_W(p->mode, k, &w, W_out + ci * (int64_t)k * k);
        } else if (W_out && lt == 0) {
          double* Wo = W_out + ci * (int64_t)k * k;
          memset(Wo, 0, sizeof(double) * (size_t)k * k);
          double f = (p->mode == ORC_MODE_REF_ETKF) ? p->inflation : sqrt(p->inflation);
          if (p->mode == ORC_MODE_REF_COMPAT) for (int i = 0; i < k; ++i) Wo[i] = f;
          else for (int i = 0; i < k; ++i) Wo[i * k + i] = f;
        }
        const int l0 = per_level ? lt : 0, l1 = per_level ? lt + 1 : nz;
        for (int lev = l0; lev < l1; ++lev) {
          const int64_t pt = (int64_t)lev * G + g;
          for (int m = 0; m < k; ++m) xk[m] = X[(int64_t)m * n + pt];
          if (have) apply_point(p->mode, p->inflation, k, &w, xk);
          else inflate_point(p->mode, p->inflation, k, xk);
          for (int m = 0; m < k; ++m) X[(int64_t)m * n + pt] = xk[m];
        }
      }
    }
    free(idx); free(idl); free(Yl); free(dl); free(rinv); ws_free(&w);
  }
  free(Yp); free(d); free(levmap);
  return rc_all;
}

/* The reference exactly as shipped (LETKF.hpp:101-113, 152-243): serial loop over grid points in
 * backend iteration order (SimpleGeometry.hpp:50-54 y-outer/x-inner; 3-D: level-major as
 * backends/wrf/WRFGeometryIterator.hpp:113-116), H re-evaluated on the LIVE, partially updated
 * ensemble (:197-206), update written in place (:240-242). */
static int letkf_as_written(const orc_letkf_params* p, double* X, const int32_t* ox,
                            const int32_t* oy, const int32_t* oz, const double* oval,
                            const double* oerr, const uint8_t* valid, int32_t* counts_out) {
  const int nx = p->nx, ny = p->ny, nz = p->nz, k = p->k;
  const int64_t P = p->P, G = (int64_t)nx * ny, n = G * nz;
  if (k > 1024) return -3;
  int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(P > 0 ? P : 1));
  double* Yl = (double*)malloc(sizeof(double) * (size_t)(P > 0 ? P : 1) * k);
  double* dl = (double*)malloc(sizeof(double) * (size_t)(P > 0 ? P : 1));
  double* rinv = (double*)malloc(sizeof(double) * (size_t)(P > 0 ? P : 1));
  xform_ws w; ws_alloc(&w, k, P);
  double xk[1024];
  int rc_all = 0;
  for (int lev = 0; lev < nz; ++lev)
    for (int gy = 0; gy < ny; ++gy)
      for (int gx = 0; gx < nx; ++gx) {
        const int64_t g = (int64_t)gy * nx + gx, pt = (int64_t)lev * G + g;
        const int64_t pl = orc_select_local(gx, gy, P, ox, oy, p->radius, idx);
        if (counts_out && lev == 0) counts_out[g] = (int32_t)pl;
        for (int m = 0; m < k; ++m) xk[m] = X[(int64_t)m * n + pt];
        if (pl == 0) {
          inflate_point(p->mode, p->inflation, k, xk);
        } else {
          for (int64_t a = 0; a < pl; ++a) {
            const int32_t i = idx[a];
            double s = 0.0;
            for (int m = 0; m < k; ++m) {
              double h = (valid && !valid[i]) ? 0.0
                         : hx_one(X + (int64_t)m * n, nx, ny, nz, ox[i], oy[i], oz ? oz[i] : 0);
              Yl[a * k + m] = h;
              s += h;
            }
            double mean = s / (double)k;
            for (int m = 0; m < k; ++m) Yl[a * k + m] -= mean;
            dl[a] = oval[i] - mean;
            double var = oerr[i] * oerr[i];
            if (valid && !valid[i]) var = INFINITY;
            rinv[a] = 1.0 / var;
          }
          int rc = local_transform(p->mode, p->inflation, k, pl, Yl, dl, rinv, &w);
          if (rc) { rc_all = rc; continue; }
          apply_point(p->mode, p->inflation, k, &w, xk);
        }
        for (int m = 0; m < k; ++m) X[(int64_t)m * n + pt] = xk[m];
      }
  free(idx); free(Yl); free(dl); free(rinv); ws_free(&w);
  return rc_all;
}

int orc_letkf(const orc_letkf_params* p, double* X, const int32_t* ox, const int32_t* oy,
              const int32_t* oz, const double* oval, const double* oerr, const uint8_t* valid,
              const int64_t* cols_sel, int64_t ncols_sel, int32_t* counts_out, double* W_out) {
  if (p->semantics == ORC_SEM_AS_WRITTEN)
    return letkf_as_written(p, X, ox, oy, oz, oval, oerr, valid, counts_out);
  return letkf_snapshot(p, NULL, X, ox, oy, oz, oval, oerr, valid, cols_sel, ncols_sel, counts_out,
                        W_out);
}

/* Snapshot LETKF with GEOGRAPHIC observation / grid locations and / or a multi-variable state (the WRF-shaped
 * case: Location.hpp:213-217 distances, IdentityObsOperator.hpp:236-281 variable access).  ox, oy, oz are the
 * observations' nearest grid points (orc_geo_locate). */
int orc_letkf_ext(const orc_letkf_params* p, const orc_ext* ext, double* X, const int32_t* ox,
                  const int32_t* oy, const int32_t* oz, const double* oval, const double* oerr,
                  const uint8_t* valid, const int64_t* cols_sel, int64_t ncols_sel, int32_t* counts_out,
                  double* W_out) {
  if (ext && ext->nvar > 63) return -4;
  return letkf_snapshot(p, ext, X, ox, oy, oz, oval, oerr, valid, cols_sel, ncols_sel, counts_out, W_out);
}

/* ------------------------------------------------------------------ global ETKF */

/* ETKF.hpp:100-179 */
int orc_etkf(double* X, int nx, int ny, int nz, int k, int64_t P, const int32_t* ox,
             const int32_t* oy, const int32_t* oz, const double* oval, const double* oerr,
             const uint8_t* valid, double inflation) {
  const int64_t n = (int64_t)nx * ny * nz;
  double* mean = (double*)malloc(sizeof(double) * (size_t)n);
  orc_ensemble_mean(X, k, n, mean);                                  /* :111 */
  double* Y = (double*)malloc(sizeof(double) * (size_t)P * k);
  double* Yp = (double*)malloc(sizeof(double) * (size_t)P * k);
  double* d = (double*)malloc(sizeof(double) * (size_t)P);
  double* rinv = (double*)malloc(sizeof(double) * (size_t)P);
  orc_obs_space(X, nx, ny, nz, k, P, ox, oy, oz, valid, oval, Y, NULL, Yp, d);   /* :128-141 */
  for (int64_t i = 0; i < P; ++i) {
    double var = (valid && !valid[i]) ? INFINITY : oerr[i] * oerr[i];           /* :144-147 */
    rinv[i] = 1.0 / var;
  }
  xform_ws w; ws_alloc(&w, k, P);
  int rc = local_transform(ORC_MODE_REF_ETKF, inflation, k, P, Yp, d, rinv, &w);  /* :150-159 */
  if (!rc) {
    double* xp = (double*)malloc(sizeof(double) * (size_t)k);
    for (int64_t pt = 0; pt < n; ++pt) {                                          /* :163-176 */
      for (int m = 0; m < k; ++m) xp[m] = (X[(int64_t)m * n + pt] - mean[pt]) * inflation; /* :125 */
      double inc = 0.0;
      for (int j = 0; j < k; ++j) inc += xp[j] * w.wa[j];
      double xam = mean[pt] + inc;
      for (int i = 0; i < k; ++i) {
        double s = 0.0;
        for (int j = 0; j < k; ++j) s += xp[j] * w.Wa[j * k + i];
        X[(int64_t)i * n + pt] = xam + s;
      }
    }
    free(xp);
  }
  ws_free(&w); free(mean); free(Y); free(Yp); free(d); free(rinv);
  return rc;
}

/* ------------------------------------------------------------------ global stochastic EnKF */

/* EnKF.hpp:139-256, 316-361 */
int orc_enkf(double* X, int nx, int ny, int nz, int k, int64_t P, const int32_t* ox,
             const int32_t* oy, const int32_t* oz, const double* oval, const double* oerr,
             const uint8_t* valid, double inflation, const double* Z, int want_gain_stats,
             int nthreads, orc_enkf_diag* diag) {
  const int64_t n = (int64_t)nx * ny * nz;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
  double* mean = (double*)malloc(sizeof(double) * (size_t)n);
  orc_ensemble_mean(X, k, n, mean);                                            /* :149 */
  const double sqi = sqrt(inflation);                                          /* :316-331 */
  double bs = 0.0;
  for (int m = 0; m < k; ++m)
    for (int64_t i = 0; i < n; ++i) {
      double v = (X[(int64_t)m * n + i] - mean[i]) * sqi;
      bs += v * v;
    }
  const double background_spread = sqrt(bs / ((double)n * k));                 /* :334 */
  double* Y = (double*)malloc(sizeof(double) * (size_t)P * k);
  double* Yp = (double*)malloc(sizeof(double) * (size_t)P * k);
  double* d = (double*)malloc(sizeof(double) * (size_t)P);
  orc_obs_space(X, nx, ny, nz, k, P, ox, oy, oz, valid, oval, Y, NULL, Yp, d); /* :168-181 */
  double dn = 0.0;
  for (int64_t i = 0; i < P; ++i) dn += d[i] * d[i];
  const double km1 = (double)(k - 1);
  /* S = Y'Y'^T/(k-1) + R  (:193-196) */
  double* S = (double*)malloc(sizeof(double) * (size_t)P * P);
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
  for (int64_t a = 0; a < P; ++a)
    for (int64_t b = 0; b < P; ++b) {
      double s = 0.0;
      for (int m = 0; m < k; ++m) s += Yp[a * k + m] * Yp[b * k + m];
      s /= km1;
      if (a == b) s += (valid && !valid[a]) ? INFINITY : oerr[a] * oerr[a];
      S[a * P + b] = s;
    }
  /* Cholesky S = L L^T (in place, lower), then solve S B = Dm for the P x k matrix
   * Dm[:,i] = yo + eps_i - Yb[:,i]  (:219-222) -- algebraically K d_i of :199,:225 */
  int rc = 0;
  double* Lm = (double*)malloc(sizeof(double) * (size_t)P * P);
  memcpy(Lm, S, sizeof(double) * (size_t)P * P);
  for (int64_t j = 0; j < P && !rc; ++j) {
    double s = Lm[j * P + j];
    for (int64_t t = 0; t < j; ++t) s -= Lm[j * P + t] * Lm[j * P + t];
    if (!(s > 0.0)) { rc = -1; break; }
    double ljj = sqrt(s);
    Lm[j * P + j] = ljj;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads) if (P - j > 256)
#endif
    for (int64_t i = j + 1; i < P; ++i) {
      double v = Lm[i * P + j];
      for (int64_t t = 0; t < j; ++t) v -= Lm[i * P + t] * Lm[j * P + t];
      Lm[i * P + j] = v / ljj;
    }
  }
  double* B = (double*)malloc(sizeof(double) * (size_t)P * k);
  if (!rc) {
    for (int64_t a = 0; a < P; ++a) {
      double sd = (valid && !valid[a]) ? INFINITY : sqrt(oerr[a] * oerr[a]); /* llt(R).matrixL() */
      for (int m = 0; m < k; ++m) B[a * k + m] = (oval[a] + sd * Z[a * k + m]) - Y[a * k + m];
    }
    for (int m = 0; m < k; ++m) {
      for (int64_t i = 0; i < P; ++i) {
        double s = B[i * k + m];
        for (int64_t t = 0; t < i; ++t) s -= Lm[i * P + t] * B[t * k + m];
        B[i * k + m] = s / Lm[i * P + i];
      }
      for (int64_t i = P - 1; i >= 0; --i) {
        double s = B[i * k + m];
        for (int64_t t = i + 1; t < P; ++t) s -= Lm[t * P + i] * B[t * k + m];
        B[i * k + m] = s / Lm[i * P + i];
      }
    }
  }
  /* C = Y'^T B / (k-1)  (k x k): xa_i = xbar + X'_i*sqi + sum_j X'_j*sqi C[j][i] */
  double* C = (double*)calloc((size_t)k * k, sizeof(double));
  if (!rc)
    for (int j = 0; j < k; ++j)
      for (int i = 0; i < k; ++i) {
        double s = 0.0;
        for (int64_t a = 0; a < P; ++a) s += Yp[a * k + j] * B[a * k + i];
        C[j * k + i] = s / km1;
      }
  double kmax = NAN, kmin = NAN, cond = NAN;
  if (!rc && want_gain_stats) {
    /* explicit S^-1 and K = X' Y'^T S^-1/(k-1) (:199-203); cond(S) via Jacobi (:206-209) */
    double* Sinv = (double*)malloc(sizeof(double) * (size_t)P * P);
    double* M = (double*)malloc(sizeof(double) * (size_t)k * P);
    orc_lu_inverse((int)P, S, Sinv);
    for (int m = 0; m < k; ++m)
      for (int64_t b = 0; b < P; ++b) {
        double s = 0.0;
        for (int64_t a = 0; a < P; ++a) s += Yp[a * k + m] * Sinv[a * P + b];
        M[m * P + b] = s;
      }
    kmax = -INFINITY; kmin = INFINITY;
    for (int64_t pt = 0; pt < n; ++pt)
      for (int64_t b = 0; b < P; ++b) {
        double s = 0.0;
        for (int m = 0; m < k; ++m) s += ((X[(int64_t)m * n + pt] - mean[pt]) * sqi) * M[m * P + b];
        s /= km1;
        if (s > kmax) kmax = s;
        if (s < kmin) kmin = s;
      }
    if (P <= 2000) {
      double* ev = (double*)malloc(sizeof(double) * (size_t)P);
      double* V = (double*)malloc(sizeof(double) * (size_t)P * P);
      int sw;
      orc_jacobi_eigh((int)P, S, ev, V, &sw);
      double emax = -INFINITY, emin = INFINITY;
      for (int64_t i = 0; i < P; ++i) { if (ev[i] > emax) emax = ev[i]; if (ev[i] < emin) emin = ev[i]; }
      cond = emax / emin;
      free(ev); free(V);
    }
    free(Sinv); free(M);
  }
  double as = 0.0;
  if (!rc) {
    double* xp = (double*)malloc(sizeof(double) * (size_t)k);
    double* xa = (double*)malloc(sizeof(double) * (size_t)k);
    for (int64_t pt = 0; pt < n; ++pt) {
      for (int m = 0; m < k; ++m) xp[m] = (X[(int64_t)m * n + pt] - mean[pt]) * sqi;
      double am = 0.0;
      for (int i = 0; i < k; ++i) {
        double s = 0.0;
        for (int j = 0; j < k; ++j) s += xp[j] * C[j * k + i];
        xa[i] = mean[pt] + (xp[i] + s);                                          /* :225-226 */
        am += xa[i];
      }
      am *= 1.0 / (double)k;
      for (int i = 0; i < k; ++i) {
        X[(int64_t)i * n + pt] = xa[i];
        as += (xa[i] - am) * (xa[i] - am);
      }
    }
    free(xp); free(xa);
  }
  if (diag) {
    diag->innovation_norm = sqrt(dn);                                           /* :184 */
    diag->background_spread = background_spread;
    diag->analysis_spread = sqrt(as / ((double)k * n));                         /* :253 */
    diag->max_kalman_gain = kmax;
    diag->min_kalman_gain = kmin;
    diag->condition_number = cond;
  }
  free(mean); free(Y); free(Yp); free(d); free(S); free(Lm); free(B); free(C);
  return rc;
}

/* ------------------------------------------------------------------ LWEnKF (locally weighted EnKF) */

/* LWEnKF.hpp:603-636: the localisation function of the NORMALISED distance */
static double lw_loc_fn(int fn, double distance, double radius) {
  const double nd = distance / radius;
  switch (fn) {
    case ORC_LOC_GAUSSIAN: return exp(-0.5 * nd * nd);
    case ORC_LOC_EXPONENTIAL: return exp(-nd);
    case ORC_LOC_CUTOFF: return nd <= 1.0 ? 1.0 : 0.0;
    case ORC_LOC_REF_GASPARI_COHN:
      if (nd >= 2.0) return 0.0;
      if (nd >= 1.0) { const double z = nd - 1.0; return ((-0.25 * z + 0.5) * z + 0.625) * z + 0.125; }
      return (((-0.25 * nd + 0.5) * nd + 0.625) * nd - 5.0) * nd + 4.0;
    default: return exp(-0.5 * nd * nd);
  }
}

/* LWEnKF<Tag>::Analyse, LWEnKF.hpp:207-334 (+ computeWeights :400-531, computeWeightedCovariance :537-551,
 * applyLocalization :556-574, applyLocalizationToGain :579-598, applyInflation :641-660,
 * generateObservationPerturbations :665-685 with the N(0,1) draws Z supplied).  Everything is global and dense, as
 * written: S is P x P, K is n x P, the "distances" of both localisations are INDEX distances |i - j| / dim.
 * X: [k][n] in place.  weighting: 0 uniform, 1 adaptive, 2 inverse_var, 3 likelihood.  diag: [9] innovation_norm,
 * background_spread, analysis_spread, max K, min K, cond(S), max w, min w, var w. */
int orc_lwenkf(double* X, int nx, int ny, int nz, int k, int64_t P, const int32_t* ox, const int32_t* oy,
               const int32_t* oz, const double* oval, const double* oerr, const uint8_t* valid, double inflation,
               double radius, int loc_fn, int weighting, const double* Z, double* diag) {
  const int64_t n = (int64_t)nx * ny * nz;
  double* mean = (double*)malloc(sizeof(double) * (size_t)n);
  orc_ensemble_mean(X, k, n, mean);                                            /* :219 */
  double* Xp = (double*)malloc(sizeof(double) * (size_t)n * k);               /* [k][n] perturbations :220-231 */
  for (int m = 0; m < k; ++m)
    for (int64_t i = 0; i < n; ++i) Xp[(int64_t)m * n + i] = X[(int64_t)m * n + i] - mean[i];
  double* Y = (double*)malloc(sizeof(double) * (size_t)P * k);
  double* Yp = (double*)malloc(sizeof(double) * (size_t)P * k);
  double* d = (double*)malloc(sizeof(double) * (size_t)P);
  orc_obs_space(X, nx, ny, nz, k, P, ox, oy, oz, valid, oval, Y, NULL, Yp, d); /* :241-255 */
  /* ---- weights (:400-436) */
  double* w = (double*)malloc(sizeof(double) * (size_t)k);
  for (int m = 0; m < k; ++m) {
    double sq = 0.0;
    for (int64_t i = 0; i < n; ++i) sq += Xp[(int64_t)m * n + i] * Xp[(int64_t)m * n + i];
    if (weighting == 0) w[m] = 1.0 / k;
    else if (weighting == 1) w[m] = 1.0 / (sqrt(sq) + 1e-8);
    else if (weighting == 2) w[m] = 1.0 / (sq + 1e-8);
    else {
      double q = 0.0;                                                          /* :514-529, R^-1 of a diagonal R */
      for (int64_t a = 0; a < P; ++a) {
        const double in = oval[a] - Y[a * k + m];
        const double var = (valid && !valid[a]) ? INFINITY : oerr[a] * oerr[a];
        q += in * ((1.0 / var) * in);
      }
      w[m] = exp(-0.5 * q);
    }
  }
  if (weighting == 1 || weighting == 2) {                                      /* :457-466 / :490-499 */
    double s = 0.0;
    for (int m = 0; m < k; ++m) s += w[m];
    for (int m = 0; m < k; ++m) w[m] /= s;
  }
  {
    double s = 0.0;                                                            /* :425-428 */
    for (int m = 0; m < k; ++m) s += w[m];
    for (int m = 0; m < k; ++m) w[m] /= s;
  }
  double wmax = w[0], wmin = w[0], wvar = 0.0;
  for (int m = 0; m < k; ++m) {
    if (w[m] > wmax) wmax = w[m];
    if (w[m] < wmin) wmin = w[m];
    const double df = w[m] - 1.0 / k;
    wvar += df * df;
  }
  wvar /= k;
  /* ---- inflation (:641-660) */
  const double sqi = sqrt(inflation);
  double bs = 0.0;
  for (int64_t e = 0; e < n * k; ++e) { Xp[e] *= sqi; bs += Xp[e] * Xp[e]; }
  const double background_spread = sqrt(bs / ((double)n * k));
  double dn = 0.0;
  for (int64_t a = 0; a < P; ++a) dn += d[a] * d[a];
  /* ---- S = (sum_m w_m y'_m y'_m^T) o L + R (:263-275) */
  double* S = (double*)malloc(sizeof(double) * (size_t)P * P);
  for (int64_t a = 0; a < P; ++a)
    for (int64_t b = 0; b < P; ++b) {
      double s = 0.0;
      for (int m = 0; m < k; ++m) s += w[m] * Yp[a * k + m] * Yp[b * k + m];
      s *= lw_loc_fn(loc_fn, fabs((double)(a - b)) / (double)P, radius);
      if (a == b) s += (valid && !valid[a]) ? INFINITY : oerr[a] * oerr[a];
      S[a * P + b] = s;
    }
  double* Sinv = (double*)malloc(sizeof(double) * (size_t)P * P);
  int rc = orc_lu_inverse((int)P, S, Sinv);                                    /* S.inverse() :279 */
  /* cond(S) from its singular values = |eigenvalues| of the symmetric S (:289-292) */
  double cond = NAN;
  if (!rc) {
    double* ev = (double*)malloc(sizeof(double) * (size_t)P);
    double* V = (double*)malloc(sizeof(double) * (size_t)P * P);
    if (!orc_jacobi_eigh((int)P, S, ev, V, NULL)) {
      double lo = INFINITY, hi = 0.0;
      for (int64_t a = 0; a < P; ++a) { const double v = fabs(ev[a]); if (v < lo) lo = v; if (v > hi) hi = v; }
      cond = hi / lo;
    }
    free(ev); free(V);
  }
  double kmax = -INFINITY, kmin = INFINITY;
  if (!rc) {
    /* K = ((X' Y'^T) S^-1) / (k - 1), then o Lg (:278-282); xa_m = xb + x'_m + K (yo + eps_m - Yb_m) (:299-310) */
    const int64_t dim = n > P ? n : P;
    double* XY = (double*)malloc(sizeof(double) * (size_t)P);
    double* Krow = (double*)malloc(sizeof(double) * (size_t)P);
    for (int64_t i = 0; i < n; ++i) {
      for (int64_t a = 0; a < P; ++a) {
        double s = 0.0;
        for (int m = 0; m < k; ++m) s += Xp[(int64_t)m * n + i] * Yp[a * k + m];
        XY[a] = s;
      }
      for (int64_t b = 0; b < P; ++b) {
        double s = 0.0;
        for (int64_t a = 0; a < P; ++a) s += XY[a] * Sinv[a * P + b];
        s = s / (double)(k - 1) * lw_loc_fn(loc_fn, fabs((double)(i - b)) / (double)dim, radius);
        Krow[b] = s;
        if (s > kmax) kmax = s;
        if (s < kmin) kmin = s;
      }
      for (int m = 0; m < k; ++m) {
        double s = 0.0;
        for (int64_t b = 0; b < P; ++b) {
          const double sd = (valid && !valid[b]) ? INFINITY : sqrt(oerr[b] * oerr[b]);
          s += Krow[b] * ((oval[b] + sd * Z[b * k + m]) - Y[b * k + m]);
        }
        X[(int64_t)m * n + i] = mean[i] + (Xp[(int64_t)m * n + i] + s);
      }
    }
    free(XY); free(Krow);
  }
  double as = 0.0;
  if (!rc) {
    orc_ensemble_mean(X, k, n, mean);                                          /* :320-333 */
    for (int m = 0; m < k; ++m)
      for (int64_t i = 0; i < n; ++i) { const double v = X[(int64_t)m * n + i] - mean[i]; as += v * v; }
  }
  if (diag) {
    diag[0] = sqrt(dn); diag[1] = background_spread; diag[2] = sqrt(as / ((double)k * n));
    diag[3] = kmax; diag[4] = kmin; diag[5] = cond; diag[6] = wmax; diag[7] = wmin; diag[8] = wvar;
  }
  free(mean); free(Xp); free(Y); free(Yp); free(d); free(w); free(S); free(Sinv);
  return rc;
}
