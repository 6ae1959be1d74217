/* cpu_baseline.c -- the CPU arm of bench.py: a performance-minded host implementation of the SAME analysis the GPU
 * path computes (canonical LETKF: Gaspari-Cohn R-localisation, symmetric square-root transform, X'W update; snapshot
 * semantics, H once), for timing on the GPU box's host cores.  TEST / BENCH INFRASTRUCTURE ONLY, like the rest of
 * oracle/ -- never on the product path.
 *
 * The reference has no multithreaded analysis and its LETKF.hpp (Eigen) cannot be built here, so the
 * "reference arm" is a port.  The parity oracle (metada_oracle.c) is written for operation-by-operation fidelity
 * (brute-force O(P) selection as LETKF.hpp:159-165, full Gram matrix, cyclic Jacobi, -O2 -ffp-contract=off) and is a
 * poor baseline: ~0.7 GFLOP/s per core.  This file is what a careful host implementation looks like without BLAS /
 * LAPACK (absent from the image):
 *   - cell index over the observations (the fair algorithmic peer of the device's bucket index),
 *   - upper-triangular SYRK with the weights folded into one operand,
 *   - Householder tridiagonalisation + implicit QL with accumulated vectors (the EISPACK tred2 / tql2 pair as
 *     published in JAMA, restated; ~9 k^3 flops instead of Jacobi's ~60 k^3),
 *   - W = B B^T with B = V diag((k-1)/lambda)^(1/4) (upper triangle), level-blocked update,
 *   - OpenMP over columns, schedule(dynamic); built -O3 -march=native with contraction on.
 * Checked against the parity oracle to 1e-10 in tests/test_cpu_baseline.py. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int nx, ny, nz, k;
  int64_t P;
  double radius, inflation;
  int nthreads;
} cpub_params;

static double gaspari_cohn(double z) {           /* Gaspari & Cohn 1999 eq. 4.10, z = dist / c, support 2c */
  z = fabs(z);
  if (z >= 2.0) return 0.0;
  if (z <= 1.0) return (((-0.25 * z + 0.5) * z + 0.625) * z - 5.0 / 3.0) * z * z + 1.0;
  const double v = ((((z / 12.0 - 0.5) * z + 0.625) * z + 5.0 / 3.0) * z - 5.0) * z + 4.0 - 2.0 / (3.0 * z);
  return v > 0.0 ? v : 0.0;
}

/* symmetric eigen-decomposition: V (row-major n x n) holds A on entry (lower triangle used), eigenvectors in its
 * columns on return; d = eigenvalues; e = scratch [n].  tred2 + tql2 (EISPACK, as restated in JAMA). */
static void tred2(int n, double* V, double* d, double* e) {
#define Vm(i, j) V[(size_t)(i) * n + (j)]
  for (int j = 0; j < n; j++) d[j] = Vm(n - 1, j);
  for (int i = n - 1; i > 0; i--) {
    double scale = 0.0, h = 0.0;
    for (int k = 0; k < i; k++) scale += fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; j++) { d[j] = Vm(i - 1, j); Vm(i, j) = 0.0; Vm(j, i) = 0.0; }
    } else {
      for (int k = 0; k < i; k++) { d[k] /= scale; h += d[k] * d[k]; }
      double f = d[i - 1], g = sqrt(h);
      if (f > 0) g = -g;
      e[i] = scale * g; h -= f * g; d[i - 1] = f - g;
      for (int j = 0; j < i; j++) e[j] = 0.0;
      for (int j = 0; j < i; j++) {
        f = d[j]; Vm(j, i) = f; g = e[j] + Vm(j, j) * f;
        for (int k = j + 1; k <= i - 1; k++) { g += Vm(k, j) * d[k]; e[k] += Vm(k, j) * f; }
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; j++) { e[j] /= h; f += e[j] * d[j]; }
      const double hh = f / (h + h);
      for (int j = 0; j < i; j++) e[j] -= hh * d[j];
      for (int j = 0; j < i; j++) {
        f = d[j]; g = e[j];
        for (int k = j; k <= i - 1; k++) Vm(k, j) -= (f * e[k] + g * d[k]);
        d[j] = Vm(i - 1, j); Vm(i, j) = 0.0;
      }
    }
    d[i] = h;
  }
  for (int i = 0; i < n - 1; i++) {
    Vm(n - 1, i) = Vm(i, i); Vm(i, i) = 1.0;
    const double h = d[i + 1];
    if (h != 0.0) {
      for (int k = 0; k <= i; k++) d[k] = Vm(k, i + 1) / h;
      for (int j = 0; j <= i; j++) {
        double g = 0.0;
        for (int k = 0; k <= i; k++) g += Vm(k, i + 1) * Vm(k, j);
        for (int k = 0; k <= i; k++) Vm(k, j) -= g * d[k];
      }
    }
    for (int k = 0; k <= i; k++) Vm(k, i + 1) = 0.0;
  }
  for (int j = 0; j < n; j++) { d[j] = Vm(n - 1, j); Vm(n - 1, j) = 0.0; }
  Vm(n - 1, n - 1) = 1.0; e[0] = 0.0;
}
static int tql2(int n, double* V, double* d, double* e) {
  for (int i = 1; i < n; i++) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  const double eps = 2.220446049250313e-16;
  for (int l = 0; l < n; l++) {
    const double t = fabs(d[l]) + fabs(e[l]);
    if (t > tst1) tst1 = t;
    int m = l;
    while (m < n) { if (fabs(e[m]) <= eps * tst1) break; m++; }
    if (m > l) {
      int iter = 0;
      do {
        if (++iter > 60) return -1;
        double g = d[l], p = (d[l + 1] - g) / (2.0 * e[l]), r = hypot(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r); d[l + 1] = e[l] * (p + r);
        const double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; i++) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1.0, c2 = c, c3 = c, s = 0.0, s2 = 0.0;
        const double el1 = e[l + 1];
        for (int i = m - 1; i >= l; i--) {
          c3 = c2; c2 = c; s2 = s;
          g = c * e[i]; h = c * p; r = hypot(p, e[i]);
          e[i + 1] = s * r; s = e[i] / r; c = p / r; p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          for (int k = 0; k < n; k++) {
            h = Vm(k, i + 1);
            Vm(k, i + 1) = s * Vm(k, i) + c * h;
            Vm(k, i) = c * Vm(k, i) - s * h;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1; e[l] = s * p; d[l] = c * p;
      } while (fabs(e[l]) > eps * tst1);
    }
    d[l] += f; e[l] = 0.0;
  }
  return 0;
#undef Vm
}

/* 4-point inverse-distance H of IdentityObsOperator.hpp:594-676 at integer observation coordinates (an exact hit
 * has weight 1e12: the value is the grid point's up to 1e-12) -- same numbers as the parity oracle's hx_one */
static double hx_point(const double* s, int nx, int ny, int nz, int oxi, int oyi, int ozi) {
  double x = oxi, y = oyi;
  int z = ozi < 0 ? 0 : (ozi >= nz ? nz - 1 : ozi);
  if (x < 0) x = 0; if (x > nx - 1) x = nx - 1;
  if (y < 0) y = 0; if (y > ny - 1) y = ny - 1;
  int i0 = (int)floor(x), j0 = (int)floor(y);
  int i1 = i0 + 1 < nx ? i0 + 1 : nx - 1, j1 = j0 + 1 < ny ? j0 + 1 : ny - 1;
  const int ii[4] = {i0, i1, i0, i1}, jj[4] = {j0, j0, j1, j1};
  double sw = 0.0, sv = 0.0;
  for (int c = 0; c < 4; ++c) {
    const double dx = x - ii[c], dy = y - jj[c], dd = sqrt(dx * dx + dy * dy);
    const double w = dd == 0.0 ? 1e12 : 1.0 / dd;
    sw += w; sv += w * s[((size_t)z * ny + jj[c]) * nx + ii[c]];
  }
  return sv / sw;
}

int cpub_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* X: [k][nz][ny][nx] analysed in place; returns 0, fills *sum_ploc.  cols_sel: optional subset of columns. */
int cpub_letkf(const cpub_params* p, double* X, const int32_t* ox, const int32_t* oy, const int32_t* oz,
               const double* oval, const double* oerr, const uint8_t* valid, const int64_t* cols_sel, int64_t ncols_sel,
               int64_t* sum_ploc) {
  const int nx = p->nx, ny = p->ny, nz = p->nz, k = p->k;
  const int64_t P = p->P, G = (int64_t)nx * ny, n = G * nz;
  const double km1 = (double)(k - 1), shift = km1 / p->inflation;
  int nthreads = p->nthreads > 0 ? p->nthreads : cpub_max_threads();
  /* ---- Y' = H(X) - mean, d = yo - mean(H(X)): once (snapshot semantics) */
  double* Yp = (double*)malloc(sizeof(double) * (size_t)(P > 0 ? P : 1) * k);
  double* dv = (double*)malloc(sizeof(double) * (size_t)(P > 0 ? P : 1));
  double* wv = (double*)malloc(sizeof(double) * (size_t)(P > 0 ? P : 1));   /* 1 / sigma^2, 0 if invalid */
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int64_t a = 0; a < P; ++a) {
    double s = 0.0;
    const int ok = valid ? valid[a] : 1;
    for (int m = 0; m < k; ++m) {
      const double y = ok ? hx_point(X + (size_t)m * n, nx, ny, nz, ox[a], oy[a], oz ? oz[a] : 0) : 0.0;
      Yp[a * k + m] = y; s += y;
    }
    s *= 1.0 / k;
    for (int m = 0; m < k; ++m) Yp[a * k + m] -= s;
    dv[a] = oval[a] - s;
    wv[a] = ok ? 1.0 / (oerr[a] * oerr[a]) : 0.0;
  }
  /* ---- cell index (cell = ceil(radius)), observations bucketed by (clamped) cell */
  const int cell = (int)ceil(p->radius) > 0 ? (int)ceil(p->radius) : 1;
  int xmin = 0, ymin = 0, xmax = nx - 1, ymax = ny - 1;
  for (int64_t a = 0; a < P; ++a) {
    if (ox[a] < xmin) xmin = ox[a]; if (ox[a] > xmax) xmax = ox[a];
    if (oy[a] < ymin) ymin = oy[a]; if (oy[a] > ymax) ymax = oy[a];
  }
  const int ncx = (xmax - xmin) / cell + 1, ncy = (ymax - ymin) / cell + 1;
  int32_t* start = (int32_t*)calloc((size_t)ncx * ncy + 1, sizeof(int32_t));
  int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)(P > 0 ? P : 1));
  for (int64_t a = 0; a < P; ++a) start[((oy[a] - ymin) / cell) * ncx + (ox[a] - xmin) / cell + 1]++;
  for (int c = 0; c < ncx * ncy; ++c) start[c + 1] += start[c];
  {
    int32_t* fill = (int32_t*)calloc((size_t)ncx * ncy, sizeof(int32_t));
    for (int64_t a = 0; a < P; ++a) {
      const int c = ((oy[a] - ymin) / cell) * ncx + (ox[a] - xmin) / cell;
      order[start[c] + fill[c]++] = (int32_t)a;
    }
    free(fill);
  }
  const int64_t ncols = cols_sel ? ncols_sel : G;
  int64_t tot_ploc = 0;
  int rc_all = 0;
  const int R = (int)floor(p->radius);
#pragma omp parallel num_threads(nthreads) reduction(+ : tot_ploc)
  {
    const size_t kk = (size_t)k * k;
    double* A = (double*)malloc(sizeof(double) * kk);
    double* B = (double*)malloc(sizeof(double) * kk);
    double* W = (double*)malloc(sizeof(double) * kk);
    double* ev = (double*)malloc(sizeof(double) * k);
    double* e = (double*)malloc(sizeof(double) * k);
    double* g = (double*)malloc(sizeof(double) * k);
    double* wbar = (double*)malloc(sizeof(double) * k);
    double* xp = (double*)malloc(sizeof(double) * k);
    int64_t cap = 256;
    double* Yw = (double*)malloc(sizeof(double) * cap * k);
    int32_t* sel = (int32_t*)malloc(sizeof(int32_t) * cap);
    double* rw = (double*)malloc(sizeof(double) * cap);
#pragma omp for schedule(dynamic, 4)
    for (int64_t ci = 0; ci < ncols; ++ci) {
      const int64_t col = cols_sel ? cols_sel[ci] : ci;
      const int gx = (int)(col % nx), gy = (int)(col / nx);
      /* selection through the cell index; ascending observation index inside the local set */
      int64_t pl = 0;
      const int cy0 = (gy - R - ymin < 0 ? 0 : gy - R - ymin) / cell, cy1 = ((gy + R - ymin) / cell < ncy - 1 ? (gy + R - ymin) / cell : ncy - 1);
      const int cx0 = (gx - R - xmin < 0 ? 0 : gx - R - xmin) / cell, cx1 = ((gx + R - xmin) / cell < ncx - 1 ? (gx + R - xmin) / cell : ncx - 1);
      for (int cy = cy0; cy <= cy1; ++cy)
        for (int32_t q = start[cy * ncx + cx0]; q < start[cy * ncx + cx1 + 1]; ++q) {
          const int32_t a = order[q];
          const double dx = (double)(gx - ox[a]), dy = (double)(gy - oy[a]), dist = sqrt(dx * dx + dy * dy);
          if (dist <= p->radius) {
            if (pl == cap) {
              cap *= 2;
              Yw = (double*)realloc(Yw, sizeof(double) * cap * k);
              sel = (int32_t*)realloc(sel, sizeof(int32_t) * cap);
              rw = (double*)realloc(rw, sizeof(double) * cap);
            }
            sel[pl] = a;
            rw[pl] = gaspari_cohn(dist / (0.5 * p->radius)) * wv[a];
            ++pl;
          }
        }
      tot_ploc += pl;
      double* xcol = X + (size_t)gy * nx + gx;              /* member m, level l at xcol[m n + l G] */
      if (pl == 0) {
        const double f = sqrt(p->inflation);
        for (int l = 0; l < nz; ++l) {
          double s = 0.0;
          for (int m = 0; m < k; ++m) s += xcol[(size_t)m * n + (size_t)l * G];
          s *= 1.0 / k;
          for (int m = 0; m < k; ++m) { double* x = &xcol[(size_t)m * n + (size_t)l * G]; *x = s + (*x - s) * f; }
        }
        continue;
      }
      /* A = shift I + Y^T diag(w) Y (lower triangle, as tred2 reads it), g = Y^T (w d) */
      for (int a = 0; a < k; ++a) g[a] = 0.0;
      memset(A, 0, sizeof(double) * kk);
      for (int64_t r = 0; r < pl; ++r) {
        const double* y = Yp + (size_t)sel[r] * k;
        double* yw = Yw + (size_t)r * k;
        const double w = rw[r], wd = w * dv[sel[r]];
        for (int a = 0; a < k; ++a) { yw[a] = w * y[a]; g[a] += y[a] * wd; }
      }
      for (int64_t r = 0; r < pl; ++r) {
        const double* y = Yp + (size_t)sel[r] * k;
        const double* yw = Yw + (size_t)r * k;
        for (int a = 0; a < k; ++a) {
          const double ya = yw[a];
          double* Ar = A + (size_t)a * k;
          for (int b = 0; b <= a; ++b) Ar[b] += ya * y[b];
        }
      }
      for (int a = 0; a < k; ++a) { A[(size_t)a * k + a] += shift; for (int b = a + 1; b < k; ++b) A[(size_t)a * k + b] = A[(size_t)b * k + a]; }
      tred2(k, A, ev, e);
      if (tql2(k, A, ev, e)) { rc_all = -2; continue; }
      /* wbar = V (V^T g / lambda);  W = B B^T, B = V diag(((k-1) / lambda)^(1/4)) */
      for (int c = 0; c < k; ++c) {
        double s = 0.0;
        for (int a = 0; a < k; ++a) s += A[(size_t)a * k + c] * g[a];
        e[c] = s / ev[c];
        xp[c] = sqrt(sqrt(km1 / ev[c]));
      }
      for (int a = 0; a < k; ++a) {
        double s = 0.0;
        for (int c = 0; c < k; ++c) { s += A[(size_t)a * k + c] * e[c]; B[(size_t)a * k + c] = A[(size_t)a * k + c] * xp[c]; }
        wbar[a] = s;
      }
      for (int a = 0; a < k; ++a)
        for (int b = 0; b <= a; ++b) {
          double s = 0.0;
          const double *Ba = B + (size_t)a * k, *Bb = B + (size_t)b * k;
          for (int c = 0; c < k; ++c) s += Ba[c] * Bb[c];
          W[(size_t)a * k + b] = s; W[(size_t)b * k + a] = s;
        }
      /* x_a = xbar + X' (wbar 1^T + W) */
      for (int l = 0; l < nz; ++l) {
        double s = 0.0, mw = 0.0;
        for (int m = 0; m < k; ++m) s += xcol[(size_t)m * n + (size_t)l * G];
        s *= 1.0 / k;
        for (int m = 0; m < k; ++m) { xp[m] = xcol[(size_t)m * n + (size_t)l * G] - s; mw += xp[m] * wbar[m]; }
        for (int i = 0; i < k; ++i) e[i] = 0.0;
        for (int j = 0; j < k; ++j) {
          const double xj = xp[j];
          const double* Wj = W + (size_t)j * k;
          for (int i = 0; i < k; ++i) e[i] += xj * Wj[i];
        }
        for (int i = 0; i < k; ++i) xcol[(size_t)i * n + (size_t)l * G] = (s + mw) + e[i];
      }
    }
    free(A); free(B); free(W); free(ev); free(e); free(g); free(wbar); free(xp); free(Yw); free(sel); free(rw);
  }
  if (sum_ploc) *sum_ploc = tot_ploc;
  free(Yp); free(dv); free(wv); free(start); free(order);
  return rc_all;
}
