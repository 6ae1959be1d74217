// C-ABI entry points for GEOGRAPHIC observations and multi-variable states; included by
// mdc_api.cu inside its extern "C" block, before the index / LETKF entry points that call geo_prepare_index().

int mdc_ens_set_geography(mdc_ens* e, const double* lat, const double* lon, int nlev, const double* vertical_coords) {
  mdc_ctx* ctx = e->ctx;
  if (!lat || !lon) MDC_FAIL(ctx, MDC_ERR_INVALID, "set_geography: null coordinate arrays");
  if (nlev < 0 || (nlev > 0 && !vertical_coords)) MDC_FAIL(ctx, MDC_ERR_INVALID, "set_geography: bad vertical coordinates");
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t G = (size_t)e->nx * e->ny;
  // extents on the host: circular mean longitude (the unwrap centre), then min / max of the unwrapped offsets
  double sx = 0.0, sy = 0.0, latmin = 1e300, latmax = -1e300, lonmin = 1e300, lonmax = -1e300;
  const double rad = 3.14159265358979323846 / 180.0;
  for (size_t i = 0; i < G; ++i) {
    if (!(lat[i] >= -90.0 && lat[i] <= 90.0) || !std::isfinite(lon[i]))
      MDC_FAIL(ctx, MDC_ERR_INVALID, "set_geography: column %zu has latitude %g longitude %g", i, lat[i], lon[i]);
    sx += std::cos(lon[i] * rad); sy += std::sin(lon[i] * rad);
    latmin = std::min(latmin, lat[i]); latmax = std::max(latmax, lat[i]);
    lonmin = std::min(lonmin, lon[i]); lonmax = std::max(lonmax, lon[i]);
  }
  const double lon_c = (sx == 0.0 && sy == 0.0) ? 0.0 : std::atan2(sy, sx) / rad;
  double umin = 1e300, umax = -1e300;
  for (size_t i = 0; i < G; ++i) {
    const double u = (lon[i] - lon_c) - 360.0 * std::rint((lon[i] - lon_c) / 360.0);
    umin = std::min(umin, u); umax = std::max(umax, u);
  }
  if (!e->glat && (dev_alloc(ctx, &e->glat, G) || dev_alloc(ctx, &e->glon, G))) return MDC_ERR_CUDA;
  MDC_CUDA(ctx, cudaMemcpyAsync(e->glat, lat, G * 8, cudaMemcpyHostToDevice, ctx->stream));
  MDC_CUDA(ctx, cudaMemcpyAsync(e->glon, lon, G * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (e->vcoord) { cudaFree(e->vcoord); e->vcoord = nullptr; }
  e->nvcoord = nlev;
  if (nlev > 0) {
    if (dev_alloc(ctx, &e->vcoord, (size_t)nlev)) return MDC_ERR_CUDA;
    MDC_CUDA(ctx, cudaMemcpyAsync(e->vcoord, vertical_coords, (size_t)nlev * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  // cells for the nearest-grid-point search: squares of the raw (lon, lat) plane holding ~4 grid points on average,
  // and a second level of GEO_COARSE x GEO_COARSE times larger ones
  {
    const double ext_x = std::max(lonmax - lonmin, 1e-9), ext_y = std::max(latmax - latmin, 1e-9);
    double c = std::max(std::sqrt(ext_x * ext_y / (double)G * 4.0), std::max(ext_x, ext_y) / 8192.0);
    while ((std::floor(ext_x / c) + 1.0) * (std::floor(ext_y / c) + 1.0) > 16777216.0) c *= 2.0;
    e->gc_lon0 = lonmin; e->gc_lat0 = latmin; e->gc_lon1 = lonmax; e->gc_lat1 = latmax;
    cudaStream_t s = ctx->stream;
    int32_t* key = nullptr;
    if (dev_alloc(ctx, &key, G)) return MDC_ERR_CUDA;
    for (int lv = 0; lv < 2; ++lv, c *= GEO_COARSE) {
      e->gc_c[lv] = c;
      e->gc_ncx[lv] = (int)std::floor(ext_x / c) + 1; e->gc_ncy[lv] = (int)std::floor(ext_y / c) + 1;
      const size_t ncell = (size_t)e->gc_ncx[lv] * e->gc_ncy[lv];
      cudaFree(e->gc_start[lv]); cudaFree(e->gc_pts[lv]); cudaFree(e->gc_plat[lv]); cudaFree(e->gc_plon[lv]);
      e->gc_start[lv] = nullptr; e->gc_pts[lv] = nullptr; e->gc_plat[lv] = nullptr; e->gc_plon[lv] = nullptr;
      int32_t* fill = nullptr;
      if (dev_alloc(ctx, &e->gc_start[lv], ncell + 1) || dev_alloc(ctx, &e->gc_pts[lv], G) || dev_alloc(ctx, &e->gc_plat[lv], G) ||
          dev_alloc(ctx, &e->gc_plon[lv], G) || dev_alloc(ctx, &fill, ncell))
        return MDC_ERR_CUDA;
      GeoCells gc{lonmin, latmin, lonmax, latmax, 1.0 / c, c, e->gc_ncx[lv], e->gc_ncy[lv], nullptr, nullptr, nullptr, nullptr};
      MDC_CUDA(ctx, cudaMemsetAsync(fill, 0, ncell * sizeof(int32_t), s));
      geo_cell_key_kernel<<<grid_for(ctx, (int64_t)G, 256, 8), 256, 0, s>>>((int64_t)G, e->glat, e->glon, gc, key, fill);
      MDC_LAUNCH_CHECK(ctx);
      index_scan_kernel<<<1, 1024, 0, s>>>(fill, e->gc_start[lv], (int)ncell);
      MDC_LAUNCH_CHECK(ctx);
      MDC_CUDA(ctx, cudaMemsetAsync(fill, 0, ncell * sizeof(int32_t), s));
      index_scatter_kernel<<<grid_for(ctx, (int64_t)G, 256, 8), 256, 0, s>>>(key, (int64_t)G, e->gc_start[lv], fill, e->gc_pts[lv]);
      MDC_LAUNCH_CHECK(ctx);
      geo_cell_gather_kernel<<<grid_for(ctx, (int64_t)G, 256, 8), 256, 0, s>>>((int64_t)G, e->gc_pts[lv], e->glat, e->glon, e->gc_plat[lv], e->gc_plon[lv]);
      MDC_LAUNCH_CHECK(ctx);
      MDC_CUDA(ctx, cudaStreamSynchronize(s));
      cudaFree(fill);
    }
    cudaFree(key);
  }
  MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  e->geo = true;
  e->geo_lon_c = lon_c; e->geo_umin = umin; e->geo_umax = umax; e->geo_latmin = latmin; e->geo_latmax = latmax;
  return MDC_OK;
}

// Geography of a decomposed ensemble = its window [gy0, gy0 + ny) x [gx0, gx0 + nx) of a store that covers the whole
// grid, with that store's frame (unwrap centre, extents): the lat / lon lattice of the bucket index is then the same
// on every rank, and so is the order in which a column meets its candidates.
int mdc_ens_set_geography_from(mdc_ens* e, const mdc_ens* g) {
  mdc_ctx* ctx = e->ctx;
  if (!g || g->ctx != ctx) MDC_FAIL(ctx, MDC_ERR_INVALID, "set_geography_from: the stores belong to different contexts");
  if (!g->geo) MDC_FAIL(ctx, MDC_ERR_INVALID, "set_geography_from: the source store has no geography");
  if (g->gnx != g->nx || g->gny != g->ny || g->nx != e->gnx || g->ny != e->gny)
    MDC_FAIL(ctx, MDC_ERR_INVALID, "set_geography_from: the source must cover the whole %d x %d grid (it is %d x %d)", e->gnx, e->gny, g->nx, g->ny);
  if (e->gx0 < 0 || e->gy0 < 0 || e->gx0 + e->nx > g->nx || e->gy0 + e->ny > g->ny)
    MDC_FAIL(ctx, MDC_ERR_INVALID, "set_geography_from: the window [%d, %d) x [%d, %d) leaves the grid", e->gx0, e->gx0 + e->nx, e->gy0, e->gy0 + e->ny);
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t G = (size_t)e->nx * e->ny_cap;
  if (!e->glat && (dev_alloc(ctx, &e->glat, G) || dev_alloc(ctx, &e->glon, G))) return MDC_ERR_CUDA;
  MDC_CUDA(ctx, cudaMemcpy2DAsync(e->glat, (size_t)e->nx * 8, g->glat + (size_t)e->gy0 * g->nx + e->gx0, (size_t)g->nx * 8,
                                  (size_t)e->nx * 8, (size_t)e->ny, cudaMemcpyDeviceToDevice, ctx->stream));
  MDC_CUDA(ctx, cudaMemcpy2DAsync(e->glon, (size_t)e->nx * 8, g->glon + (size_t)e->gy0 * g->nx + e->gx0, (size_t)g->nx * 8,
                                  (size_t)e->nx * 8, (size_t)e->ny, cudaMemcpyDeviceToDevice, ctx->stream));
  if (e->vcoord) { cudaFree(e->vcoord); e->vcoord = nullptr; }
  e->nvcoord = g->nvcoord;
  if (g->nvcoord > 0) {
    if (dev_alloc(ctx, &e->vcoord, (size_t)g->nvcoord)) return MDC_ERR_CUDA;
    MDC_CUDA(ctx, cudaMemcpyAsync(e->vcoord, g->vcoord, (size_t)g->nvcoord * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  e->geo = true;
  e->geo_lon_c = g->geo_lon_c; e->geo_umin = g->geo_umin; e->geo_umax = g->geo_umax;
  e->geo_latmin = g->geo_latmin; e->geo_latmax = g->geo_latmax;
  return MDC_OK;
}

/* the frame of a store's geography: unwrap centre of the longitudes and the extents of the unwrapped offsets /
 * latitudes over its columns -- what a caller needs to decide which observations a rank's columns can reach */
int mdc_ens_geography_frame(const mdc_ens* e, double* lon_c, double* umin, double* umax, double* latmin, double* latmax) {
  if (!e->geo) MDC_FAIL(e->ctx, MDC_ERR_INVALID, "geography_frame: the store has no geography");
  if (lon_c) *lon_c = e->geo_lon_c;
  if (umin) *umin = e->geo_umin;
  if (umax) *umax = e->geo_umax;
  if (latmin) *latmin = e->geo_latmin;
  if (latmax) *latmax = e->geo_latmax;
  return MDC_OK;
}

int mdc_ens_set_variables(mdc_ens* e, int nvar, const int32_t* var_nlev) {
  mdc_ctx* ctx = e->ctx;
  if (nvar < 1 || nvar > MDC_MAX_VARS || !var_nlev) MDC_FAIL(ctx, MDC_ERR_INVALID, "set_variables: 1 <= nvar <= %d", MDC_MAX_VARS);
  // The geometry's level count nzg is that of the mass-level variables = the smallest multi-level count; a variable
  // staggered in the vertical (WRF's W, PH, PHB: WRFState.hpp:89-93, 445-449) has nzg + 1.  H searches its neighbours
  // on the geometry's nzg levels and reads var[kk][jj][ii] with the variable's own dimensions
  // (IdentityObsOperator.hpp:684-711), so it never touches a staggered variable's top level; the column update
  // transforms every level of every variable.
  int total = 0, nzg = 0;
  for (int v = 0; v < nvar; ++v) {
    if (var_nlev[v] < 1) MDC_FAIL(ctx, MDC_ERR_INVALID, "set_variables: variable %d has %d levels", v, var_nlev[v]);
    total += var_nlev[v];
    if (var_nlev[v] > 1) nzg = nzg ? std::min(nzg, (int)var_nlev[v]) : (int)var_nlev[v];
  }
  if (!nzg) nzg = 1;
  if (total != e->nz) MDC_FAIL(ctx, MDC_ERR_INVALID, "set_variables: levels sum to %d, the ensemble has %d", total, e->nz);
  if (e->nvcoord > 0 && nzg > 1 && e->nvcoord != nzg)
    MDC_FAIL(ctx, MDC_ERR_INVALID, "set_variables: the geometry has %d vertical coordinates, the mass-level variables %d levels", e->nvcoord, nzg);
  for (int v = 0; v < nvar; ++v)
    if (var_nlev[v] != 1 && var_nlev[v] != nzg && var_nlev[v] != nzg + 1)
      MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "set_variables: variable %d has %d levels; a 3-D variable has the geometry's %d levels or one more (vertically staggered)", v, var_nlev[v], nzg);
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<int32_t> map((size_t)e->nz);
  int off = 0;
  for (int v = 0; v < nvar; ++v) {
    e->var_off[v] = off; e->var_nlev[v] = var_nlev[v];
    for (int l = 0; l < var_nlev[v]; ++l) map[(size_t)off + l] = l;
    off += var_nlev[v];
  }
  if (!e->levmap && dev_alloc(ctx, &e->levmap, (size_t)e->nz)) return MDC_ERR_CUDA;
  MDC_CUDA(ctx, cudaMemcpyAsync(e->levmap, map.data(), (size_t)e->nz * 4, cudaMemcpyHostToDevice, ctx->stream));
  MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  e->nvar = nvar; e->nzg = nzg;
  return MDC_OK;
}

int mdc_obs_set_variables(mdc_obs* o, const int32_t* var) {
  mdc_ctx* ctx = o->ctx;
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (o->P != o->P_own) MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "obs_set_variables: not with halo rows");
  if (o->var) { cudaFree(o->var); o->var = nullptr; }
  o->var_max = 0;
  o->have_hx = false;
  if (!var || o->P == 0) return MDC_OK;
  int vmax = 0;
  for (int64_t i = 0; i < o->P; ++i) {
    if (var[i] < 0 || var[i] >= MDC_MAX_VARS) MDC_FAIL(ctx, MDC_ERR_INVALID, "obs_set_variables: observation %lld observes variable %d", (long long)i, var[i]);
    vmax = std::max(vmax, (int)var[i]);
  }
  if (dev_alloc(ctx, &o->var, (size_t)o->P)) return MDC_ERR_CUDA;
  o->var_cap = (size_t)o->P;
  MDC_CUDA(ctx, cudaMemcpyAsync(o->var, var, (size_t)o->P * 4, cudaMemcpyHostToDevice, ctx->stream));
  MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  o->var_max = vmax;
  return MDC_OK;
}

int mdc_obs_create_geographic(mdc_ctx* ctx, int64_t P, const double* lat, const double* lon, const double* level,
                              const double* value, const double* err, const uint8_t* valid, const int64_t* gid,
                              mdc_obs** out) {
  if (!ctx || !out) return MDC_ERR_INVALID;
  *out = nullptr;
  if (P > 0 && (!lat || !lon)) MDC_FAIL(ctx, MDC_ERR_INVALID, "mdc_obs_create_geographic: null coordinates");
  for (int64_t i = 0; i < P; ++i)
    if (!(lat[i] >= -90.0 && lat[i] <= 90.0) || !std::isfinite(lon[i]))
      MDC_FAIL(ctx, MDC_ERR_INVALID, "mdc_obs_create_geographic: observation %lld has latitude %g longitude %g", (long long)i, lat[i], lon[i]);
  // the grid coordinates are filled by mdc_obs_locate(); create the store with zeros
  std::vector<int32_t> zero((size_t)std::max<int64_t>(P, 1), 0);
  mdc_obs* o = nullptr;
  if (int rc = mdc_obs_create(ctx, P, zero.data(), zero.data(), zero.data(), value, err, valid, gid, &o)) return rc;
  o->geo = true;
  const size_t n = (size_t)std::max<int64_t>(P, 1);
  if (dev_alloc(ctx, &o->lat, n) || dev_alloc(ctx, &o->lon, n) || dev_alloc(ctx, &o->lev, n) ||
      dev_alloc(ctx, &o->qx, n) || dev_alloc(ctx, &o->qy, n)) { mdc_obs_destroy(o); return MDC_ERR_CUDA; }
  o->geo_cap = n;
  if (P > 0) {
    cudaStream_t s = ctx->stream;
    MDC_CUDA(ctx, cudaMemcpyAsync(o->lat, lat, P * 8, cudaMemcpyHostToDevice, s));
    MDC_CUDA(ctx, cudaMemcpyAsync(o->lon, lon, P * 8, cudaMemcpyHostToDevice, s));
    if (level) MDC_CUDA(ctx, cudaMemcpyAsync(o->lev, level, P * 8, cudaMemcpyHostToDevice, s));
    else MDC_CUDA(ctx, cudaMemsetAsync(o->lev, 0, P * 8, s));
    MDC_CUDA(ctx, cudaStreamSynchronize(s));
  }
  *out = o;
  return MDC_OK;
}

int mdc_obs_locate(mdc_obs* o, mdc_ens* e) {
  mdc_ctx* ctx = o->ctx;
  if (e->ctx != ctx) MDC_FAIL(ctx, MDC_ERR_INVALID, "obs_locate: ens/obs belong to different contexts");
  if (!o->geo) MDC_FAIL(ctx, MDC_ERR_INVALID, "obs_locate: the observations carry GRID coordinates already");
  if (!e->geo) MDC_FAIL(ctx, MDC_ERR_INVALID, "obs_locate: the ensemble has no geography (mdc_ens_set_geography)");
  if (e->gnx != e->nx || e->gny != e->ny)
    MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "obs_locate: the nearest grid point is searched on the WHOLE grid: locate on an ensemble store that covers it (a 1-level, 1-member store with the global geography will do), then use mdc_ens_set_geography_from on the decomposed one");
  if (!e->gc_start[0])
    MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "obs_locate: this ensemble's geography is a window of another store's (mdc_ens_set_geography_from): locate on that one");
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (o->P > 0 && getenv("MDC_GEO_LOCATE_BRUTE")) {   // the reference's O(P G) scan, kept as a cross-check
    geo_locate_kernel<<<mdc_div_up(o->P, 256), 256, 0, ctx->stream>>>(o->P, o->lat, o->lon, o->lev, e->glat, e->glon,
                                                                    (int64_t)e->nx * e->ny, e->nx, e->vcoord, e->nvcoord,
                                                                    o->x, o->y, o->z);
    MDC_LAUNCH_CHECK(ctx);
  } else if (o->P > 0) {                                // ring walk over the bucketed grid points, same result
    GeoCells lv[2];
    for (int l = 0; l < 2; ++l)
      lv[l] = GeoCells{e->gc_lon0, e->gc_lat0, e->gc_lon1, e->gc_lat1, 1.0 / e->gc_c[l], e->gc_c[l], e->gc_ncx[l], e->gc_ncy[l],
                       e->gc_start[l], e->gc_pts[l], e->gc_plat[l], e->gc_plon[l]};
    geo_locate_ring_kernel<<<mdc_div_up(o->P, 128), 128, 0, ctx->stream>>>(o->P, o->lat, o->lon, o->lev, lv[0], lv[1], e->nx,
                                                                         e->vcoord, e->nvcoord, o->x, o->y, o->z);
    MDC_LAUNCH_CHECK(ctx);
  }
  o->located = true;
  o->have_hx = false;
  o->index_valid = false;   // the index carries the located levels
  return MDC_OK;
}

int mdc_obs_download_grid_coords(mdc_obs* o, int32_t* x, int32_t* y, int32_t* z) {
  mdc_ctx* ctx = o->ctx;
  if (o->geo && !o->located) MDC_FAIL(ctx, MDC_ERR_INVALID, "obs_download_grid_coords: call mdc_obs_locate first");
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t P = (size_t)o->P;
  if (x) MDC_CUDA(ctx, cudaMemcpyAsync(x, o->x, P * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (y) MDC_CUDA(ctx, cudaMemcpyAsync(y, o->y, P * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (z) MDC_CUDA(ctx, cudaMemcpyAsync(z, o->z, P * 4, cudaMemcpyDeviceToHost, ctx->stream));
  MDC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MDC_OK;
}

static int index_build_impl(mdc_obs* o, int cell);

// Lattice + bucket index for a haversine selection of `radius` kilometres around the columns of `e`.
static int geo_prepare_index(mdc_obs* o, mdc_ens* e, double radius) {
  mdc_ctx* ctx = o->ctx;
  if (!e->geo) MDC_FAIL(ctx, MDC_ERR_INVALID, "geographic observations need an ensemble with geography (mdc_ens_set_geography)");
  // (a decomposed domain: the ensemble's geography is a window of the global one and carries the GLOBAL frame --
  // mdc_ens_set_geography_from --, so the lattice, hence every column's candidate order, is that of the one-shot run)
  MDC_CUDA(ctx, cudaSetDevice(ctx->device));
  const double pi = 3.14159265358979323846;
  // An observation within `radius` km (great circle, R = 6371 km, Location.hpp:325) of a column at latitude phi
  // differs from it by at most delta = radius / R in latitude and asin(sin delta / cos phi) in longitude.
  const double delta = std::max(radius, 0.0) / 6371.0;
  const double phic = std::max(std::fabs(e->geo_latmin), std::fabs(e->geo_latmax)) * pi / 180.0;
  if (!(phic + delta < 0.5 * pi * (1.0 - 1e-6)))
    MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "geographic selection: a %g km circle around latitude %g reaches a pole", radius, phic * 180.0 / pi);
  const double dlat = delta * 180.0 / pi * (1.0 + 1e-9) + 1e-12;
  const double dlon = std::asin(std::min(1.0, std::sin(delta) / std::cos(phic))) * 180.0 / pi * (1.0 + 1e-9) + 1e-12;
  if (!(e->geo_umax + dlon < 180.0 && e->geo_umin - dlon > -180.0))
    MDC_FAIL(ctx, MDC_ERR_UNSUPPORTED, "geographic selection: the domain (+ radius) spans the whole longitude circle; regional domains only");
  const int R = GEO_SUB + 2;
  GeoLattice g;
  g.lon_c = e->geo_lon_c; g.u0 = e->geo_umin; g.lat0 = e->geo_latmin;
  const double ext_x = e->geo_umax - e->geo_umin, ext_y = e->geo_latmax - e->geo_latmin;
  // no finer than 1/8192 of the domain: bounds the cell count for a radius that is tiny against the domain (a coarser
  // lattice only adds candidates; the haversine test decides)
  const double qx = std::max({dlon / GEO_SUB, ext_x / 8192.0, 1e-9}), qy = std::max({dlat / GEO_SUB, ext_y / 8192.0, 1e-9});
  g.inv_qx = 1.0 / qx; g.inv_qy = 1.0 / qy;
  // columns sit in [-1, floor(ext / q) + 1] (rounding); an observation inside the radius is within R of its column
  g.lo_x = -(double)(R + 1); g.hi_x = std::floor(ext_x * g.inv_qx) + (double)(R + 2);
  g.lo_y = -(double)(R + 1); g.hi_y = std::floor(ext_y * g.inv_qy) + (double)(R + 2);
  const size_t G = (size_t)e->nx * e->ny;
  if (G > o->cq_cap) {
    cudaFree(o->cqx); cudaFree(o->cqy);
    o->cq_cap = 0;
    if (dev_alloc(ctx, &o->cqx, G) || dev_alloc(ctx, &o->cqy, G)) return MDC_ERR_CUDA;
    o->cq_cap = G;
  }
  geo_quantise_kernel<<<grid_for(ctx, (int64_t)G, 256, 8), 256, 0, ctx->stream>>>((int64_t)G, e->glat, e->glon, g, o->cqx, o->cqy);
  MDC_LAUNCH_CHECK(ctx);
  if (o->P > 0) {
    geo_quantise_kernel<<<grid_for(ctx, o->P, 256, 8), 256, 0, ctx->stream>>>(o->P, o->lat, o->lon, g, o->qx, o->qy);
    MDC_LAUNCH_CHECK(ctx);
  }
  if (int rc = index_build_impl(o, R)) return rc;
  if ((size_t)o->P > o->sgeo_cap) {
    cudaFree(o->slat); cudaFree(o->slon);
    o->sgeo_cap = 0;
    if (dev_alloc(ctx, &o->slat, (size_t)o->P) || dev_alloc(ctx, &o->slon, (size_t)o->P)) return MDC_ERR_CUDA;
    o->sgeo_cap = (size_t)o->P;
  }
  if (o->P > 0) {
    geo_gather_sorted_kernel<<<grid_for(ctx, o->P, 256, 8), 256, 0, ctx->stream>>>(o->P, o->sorted_row, o->lat, o->lon, o->slat, o->slon);
    MDC_LAUNCH_CHECK(ctx);
  }
  o->geo_reach = R; o->geo_radius = radius; o->geo_ens = e;
  return MDC_OK;
}
