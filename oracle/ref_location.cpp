// Runs the reference's own Location::distance_to (framework/base/Location.hpp, header-only, unmodified) on
// GEOGRAPHIC pairs from a binary input file and dumps the distances -- used by tests/golden/make_goldens.py
// to pin orc_distance_geo.  Test infrastructure only.
//   in : int64 n, then n x 4 doubles (lat1, lon1, lat2, lon2)
//   out: n doubles (kilometres)
// With a third argument "cartesian": n x 6 doubles (x1, y1, z1, x2, y2, z2) as CARTESIAN locations
// (Location.hpp:217-225, 3-D Euclid) -- pins orc_distance_cartesian.
#include <cstdint>
#include <cstdio>
#include <vector>

#include "Location.hpp"

int main(int argc, char** argv) {
  if (argc != 3 && argc != 4) { std::fprintf(stderr, "usage: ref_location in.bin out.bin [cartesian]\n"); return 2; }
  const bool cartesian = argc == 4;
  const int w = cartesian ? 6 : 4;
  std::FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 3;
  int64_t n = 0;
  if (std::fread(&n, 8, 1, f) != 1) return 4;
  std::vector<double> in((size_t)n * w), out((size_t)n);
  if (std::fread(in.data(), 8, in.size(), f) != in.size()) return 5;
  std::fclose(f);
  using metada::framework::CoordinateSystem;
  using metada::framework::Location;
  for (int64_t i = 0; cartesian && i < n; ++i) {
    const Location a(in[6 * i], in[6 * i + 1], in[6 * i + 2], CoordinateSystem::CARTESIAN);
    const Location b(in[6 * i + 3], in[6 * i + 4], in[6 * i + 5], CoordinateSystem::CARTESIAN);
    out[(size_t)i] = a.distance_to(b);
  }
  for (int64_t i = 0; !cartesian && i < n; ++i) {
    const Location a(in[4 * i], in[4 * i + 1], 0.0, CoordinateSystem::GEOGRAPHIC);
    const Location b(in[4 * i + 2], in[4 * i + 3], 850.0, CoordinateSystem::GEOGRAPHIC);   // the level is ignored
    out[(size_t)i] = a.distance_to(b);
  }
  std::FILE* o = std::fopen(argv[2], "wb");
  std::fwrite(out.data(), 8, out.size(), o);
  std::fclose(o);
  return 0;
}
