// Runs the reference's own Location::distance_to (framework/base/Location.hpp, header-only, unmodified) on
// GEOGRAPHIC pairs from a binary input file and dumps the distances -- used by tests/golden/make_goldens.py
// to pin orc_distance_geo.  Test infrastructure only.
//   in : int64 n, then n x 4 doubles (lat1, lon1, lat2, lon2)
//   out: n doubles (kilometres)
#include <cstdint>
#include <cstdio>
#include <vector>

#include "Location.hpp"

int main(int argc, char** argv) {
  if (argc != 3) { std::fprintf(stderr, "usage: ref_location in.bin out.bin\n"); return 2; }
  std::FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 3;
  int64_t n = 0;
  if (std::fread(&n, 8, 1, f) != 1) return 4;
  std::vector<double> in((size_t)n * 4), out((size_t)n);
  if (std::fread(in.data(), 8, in.size(), f) != in.size()) return 5;
  std::fclose(f);
  using metada::framework::CoordinateSystem;
  using metada::framework::Location;
  for (int64_t i = 0; i < n; ++i) {
    const Location a(in[4 * i], in[4 * i + 1], 0.0, CoordinateSystem::GEOGRAPHIC);
    const Location b(in[4 * i + 2], in[4 * i + 3], 850.0, CoordinateSystem::GEOGRAPHIC);   // the level is ignored
    out[(size_t)i] = a.distance_to(b);
  }
  std::FILE* o = std::fopen(argv[2], "wb");
  std::fwrite(out.data(), 8, out.size(), o);
  std::fclose(o);
  return 0;
}
