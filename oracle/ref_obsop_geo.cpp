// Runs the reference's own IdentityObsOperator (backends/common/obsoperator/IdentityObsOperator.hpp, header-only,
// unmodified) on GEOGRAPHIC observations of a multi-variable state held by mock backends that expose exactly what the
// operator asks a WRF-type backend for: geometry().unstaggered_info() with 2-D coordinate arrays and vertical
// coordinates (:488-526), x_dim / y_dim / z_dim (:684-689), state.at(name, index) and getVariableDimensions(name)
// (:694-711).  Used by tests/golden/make_goldens.py to pin orc_geo_locate + the oracle's per-variable H.
// Test infrastructure only.
//   in : int64 nx, ny, nz, nvar, P; int64 var_nlev[nvar]; lat[ny*nx], lon[ny*nx], vc[nz]; state[sum(var_nlev)][ny][nx];
//        olat[P], olon[P], olev[P], ovar[P] (double), valid[P] (double)
//   out: P doubles H(x)
#include <algorithm>   // (the reference header uses std::sort without including it)
#include <cstdint>
#include <cstdio>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "IdentityObsOperator.hpp"

namespace fwk = metada::framework;

struct GridInfo {
  size_t nx = 0, ny = 0;
  std::vector<double> longitude_2d, latitude_2d, vertical_coords;
  bool has_2d_coords() const { return !longitude_2d.empty(); }
  bool has_vertical_coords() const { return !vertical_coords.empty(); }
};
struct MockGeometry {
  GridInfo info;
  size_t nz = 1;
  const GridInfo& unstaggered_info() const { return info; }
  size_t x_dim() const { return info.nx; }
  size_t y_dim() const { return info.ny; }
  size_t z_dim() const { return nz; }
  // only reached by the operator's fallback for geometries WITHOUT coordinate arrays (:532-586); compiled, never run
  fwk::Location getLocation(int i, int j, int k) const {
    return fwk::Location(info.latitude_2d[(size_t)j * info.nx + i], info.longitude_2d[(size_t)j * info.nx + i],
                         info.vertical_coords.empty() ? 0.0 : info.vertical_coords[(size_t)k], fwk::CoordinateSystem::GEOGRAPHIC);
  }
};
struct MockState {
  const MockGeometry* geo;
  std::map<std::string, std::vector<double>> vars;
  std::map<std::string, std::vector<size_t>> dims;
  const MockGeometry& geometry() const { return *geo; }
  const double& at(const std::string& name, size_t idx) const { return vars.at(name).at(idx); }
  const std::vector<size_t>& getVariableDimensions(const std::string& name) const { return dims.at(name); }
};
struct ObsPoint {
  fwk::Location location;
  double value, error;
  bool is_valid;
};
struct MockObs {
  using value_type = ObsPoint;
  std::vector<ObsPoint> pts;
  auto begin() const { return pts.begin(); }
  auto end() const { return pts.end(); }
  size_t size() const { return pts.size(); }
};
struct MockValue {
  std::vector<std::string> v;
  std::vector<std::string> asVectorString() const { if (v.empty()) throw std::runtime_error("unset"); return v; }
};
struct MockConfig {
  std::string var;
  MockValue Get(const std::string& key) const { return key == "required_state_vars" ? MockValue{{var}} : MockValue{}; }
};
struct NoControl {};

static std::vector<double> rd(std::FILE* f, size_t n) {
  std::vector<double> v(n);
  if (n && std::fread(v.data(), 8, n, f) != n) throw std::runtime_error("short input");
  return v;
}

int main(int argc, char** argv) {
  if (argc != 3) { std::fprintf(stderr, "usage: ref_obsop_geo in.bin out.bin\n"); return 2; }
  std::FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 3;
  int64_t h[5];
  if (std::fread(h, 8, 5, f) != 5) return 4;
  const size_t nx = h[0], ny = h[1], nz = h[2], nvar = h[3], P = h[4], G = nx * ny;
  std::vector<int64_t> nlev(nvar);
  if (std::fread(nlev.data(), 8, nvar, f) != nvar) return 5;
  MockGeometry geo;
  geo.info.nx = nx; geo.info.ny = ny; geo.nz = nz;
  geo.info.latitude_2d = rd(f, G); geo.info.longitude_2d = rd(f, G); geo.info.vertical_coords = rd(f, nz);
  MockState state{&geo, {}, {}};
  for (size_t v = 0; v < nvar; ++v) {
    const std::string name = "v" + std::to_string(v);
    state.vars[name] = rd(f, (size_t)nlev[v] * G);
    state.dims[name] = nlev[v] > 1 ? std::vector<size_t>{(size_t)nlev[v], ny, nx} : std::vector<size_t>{ny, nx};
  }
  auto olat = rd(f, P), olon = rd(f, P), olev = rd(f, P), ovar = rd(f, P), valid = rd(f, P);
  std::fclose(f);
  MockObs obs;
  for (size_t i = 0; i < P; ++i)
    obs.pts.push_back({fwk::Location(olat[i], olon[i], olev[i], fwk::CoordinateSystem::GEOGRAPHIC), 0.0, 1.0, valid[i] != 0.0});
  std::vector<double> out(P, 0.0);
  using Op = metada::backends::common::obsoperator::IdentityObsOperator<MockState, MockObs, NoControl>;
  for (size_t v = 0; v < nvar; ++v) {
    Op op(MockConfig{"v" + std::to_string(v)});
    const std::vector<double> hx = op.apply(state, obs);          // the reference's H for variable v at every observation
    for (size_t i = 0; i < P; ++i) if ((size_t)ovar[i] == v) out[i] = hx[i];
  }
  std::FILE* o = std::fopen(argv[2], "wb");
  std::fwrite(out.data(), 8, P, o);
  std::fclose(o);
  return 0;
}
