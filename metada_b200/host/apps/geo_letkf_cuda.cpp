// geo_letkf_cuda <in.bin> <out.bin>: the host layer's GEOGRAPHIC path end to end -- observations carrying the
// reference's Location(lat, lon, level, GEOGRAPHIC) (framework/base/Location.hpp:82-84) go through
// DeviceObservations (CudaApi.hpp), the grid's 2-D coordinate arrays through DeviceEnsemble::setGeography, the
// analysis through mdc_letkf_analyse.  Used by tests/test_host_drivers.py; a WRF-type geometry backend reaches the
// same calls through DeviceAnalysis.hpp::uploadEnsemble (unstaggered_info()).
//   in : int64 nx, ny, nz, k, P, nvc; double radius_km; lat[ny*nx], lon[ny*nx], vc[nvc]; X[k][nz][ny][nx];
//        olat[P], olon[P], olev[P], value[P], error[P], valid[P] (as doubles)
//   out: X_a [k][nz][ny][nx]
// With MDC_WORLD_SIZE > 1 (one process per GPU: MDC_RANK, MDC_COMM_ID_FILE, MDC_DEVICE) or MDC_GEO_SHARDED=1 the
// analysis goes through the C++ runtime's sharded geographic path (mdc_geo_sharded_analyse: global geography for
// locating, a window of it per rank, halo rows by box over NCCL) and every rank writes the whole analysis.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "CudaApi.hpp"
#include "Location.hpp"

namespace fwk = metada::framework;
namespace cuda = metada::backends::cuda;

struct Point {           // what an observation backend's iterator yields (PointObservation.hpp:63-67)
  fwk::Location location;
  double value, error;
  bool is_valid;
};

static std::vector<double> rd(std::FILE* f, size_t n) {
  std::vector<double> v(n);
  if (n && std::fread(v.data(), 8, n, f) != n) throw std::runtime_error("short input file");
  return v;
}

int main(int argc, char** argv) {
  if (argc != 3) { std::fprintf(stderr, "usage: geo_letkf_cuda in.bin out.bin\n"); return 2; }
  try {
    std::FILE* f = std::fopen(argv[1], "rb");
    if (!f) throw std::runtime_error("cannot open input");
    int64_t h[6];
    double radius;
    if (std::fread(h, 8, 6, f) != 6 || std::fread(&radius, 8, 1, f) != 1) throw std::runtime_error("short header");
    const int nx = (int)h[0], ny = (int)h[1], nz = (int)h[2], k = (int)h[3];
    const size_t P = (size_t)h[4], nvc = (size_t)h[5], G = (size_t)nx * ny, n = G * nz;
    auto lat = rd(f, G), lon = rd(f, G), vc = rd(f, nvc), X = rd(f, n * k);
    auto olat = rd(f, P), olon = rd(f, P), olev = rd(f, P), val = rd(f, P), err = rd(f, P), valid = rd(f, P);
    std::fclose(f);
    std::vector<Point> obs;
    for (size_t i = 0; i < P; ++i)
      obs.push_back({fwk::Location(olat[i], olon[i], olev[i], fwk::CoordinateSystem::GEOGRAPHIC), val[i], err[i], valid[i] != 0.0});
    std::vector<const double*> in;
    std::vector<double*> out;
    for (int m = 0; m < k; ++m) { in.push_back(X.data() + (size_t)m * n); out.push_back(X.data() + (size_t)m * n); }
    const cuda::ProcessGroup pg;
    if (pg.world > 1 || std::getenv("MDC_GEO_SHARDED")) {
      mdc_stream_config cfg{};
      cfg.gnx = nx; cfg.gny = ny; cfg.nz = nz; cfg.k = k; cfg.row0 = pg.row0(ny); cfg.row1 = pg.row1(ny); cfg.slots = 3;
      mdc_stream* st = nullptr;
      if (mdc_stream_create(pg.device, &cfg, &st)) throw std::runtime_error("mdc_stream_create failed");
      auto fail = [&](const std::string& what) {
        const std::string msg = what + ": " + mdc_stream_last_error(st);
        mdc_stream_destroy(st);
        throw std::runtime_error(msg);
      };
      try { pg.attach(st); } catch (const std::exception& e) { fail(e.what()); }
      std::vector<uint8_t> ok(P);
      for (size_t i = 0; i < P; ++i) ok[i] = valid[i] != 0.0;
      mdc_letkf_params p{};
      p.radius = radius; p.inflation = 1.0; p.mode = MDC_MODE_CANONICAL; p.loc = MDC_LOC_GASPARI_COHN; p.use_R = 1;
      mdc_letkf_stats st_{};
      if (mdc_geo_sharded_analyse(st, out.data(), lat.data(), lon.data(), (int)nvc, vc.data(), 0, nullptr, (int64_t)P, olat.data(),
                                  olon.data(), olev.data(), val.data(), err.data(), ok.data(), nullptr, &p, &st_))
        fail("mdc_geo_sharded_analyse");
      if (pg.world > 1 && mdc_comm_allgather_rows(st, out.data())) fail("mdc_comm_allgather_rows");
      mdc_stream_destroy(st);
      std::FILE* o = std::fopen(argv[2], "wb");
      std::fwrite(X.data(), 8, X.size(), o);
      std::fclose(o);
      std::cout << "rank " << pg.rank << " of " << pg.world << ": columns " << st_.columns << std::endl;
      return 0;
    }
    cuda::DeviceEnsemble ens(nx, ny, nz, k);
    ens.upload(in);
    ens.setGeography(lat, lon, vc);
    cuda::DeviceObservations dobs(obs);
    if (!dobs.geographic()) throw std::runtime_error("observations were not recognised as GEOGRAPHIC");
    mdc_letkf_params p{};
    p.radius = radius; p.inflation = 1.0; p.mode = MDC_MODE_CANONICAL; p.loc = MDC_LOC_GASPARI_COHN; p.use_R = 1;
    mdc_letkf_stats st{};
    cuda::DeviceContext::Instance().check(mdc_letkf_analyse(ens.get(), dobs.get(), &p, &st), "mdc_letkf_analyse");
    ens.download(out);
    std::FILE* o = std::fopen(argv[2], "wb");
    std::fwrite(X.data(), 8, X.size(), o);
    std::fclose(o);
    std::cout << "columns " << st.columns << " mean local obs " << (double)st.sum_local_obs / (double)st.columns << std::endl;
    // mixing coordinate systems must throw, as Location::distance_to does (Location.hpp:226-229)
    obs.push_back({fwk::Location(1, 2, 0), 0.0, 1.0, true});
    try { cuda::DeviceObservations bad(obs); std::cerr << "mixed systems accepted\n"; return 4; } catch (const std::runtime_error&) {}
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "geo_letkf_cuda: " << e.what() << std::endl;
    return 1;
  }
}
