// Runs the reference's own Metrics.hpp (framework/algorithms/Metrics.hpp, header-only, unmodified)
// on a binary input file and dumps its results -- used by tests/golden/make_goldens.py to pin
// orc_metrics.  Test infrastructure only.
//   in : int64 k, int64 n, k*n doubles (member major), n doubles (truth)
//   out: 5 doubles (rmse, bias, correlation, crps, avg_spread), n doubles mean, n doubles spread
#include <cstdint>
#include <cstdio>
#include <iomanip>
#include <sstream>
#include <string>
#include <vector>
#include <cmath>

#include "Metrics.hpp"

int main(int argc, char** argv) {
  if (argc != 3) { std::fprintf(stderr, "usage: ref_metrics in.bin out.bin\n"); return 2; }
  std::FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 3;
  int64_t k = 0, n = 0;
  if (std::fread(&k, 8, 1, f) != 1 || std::fread(&n, 8, 1, f) != 1) return 4;
  std::vector<std::vector<double>> ens((size_t)k, std::vector<double>((size_t)n));
  for (auto& m : ens)
    if (std::fread(m.data(), 8, (size_t)n, f) != (size_t)n) return 5;
  std::vector<double> truth((size_t)n);
  if (std::fread(truth.data(), 8, (size_t)n, f) != (size_t)n) return 6;
  std::fclose(f);
  auto v = metada::framework::Metrics<double>::CalculateAll(ens, truth, (size_t)n, (size_t)k);
  std::FILE* o = std::fopen(argv[2], "wb");
  const double s[5] = {v.rmse, v.bias, v.correlation, v.crps, v.avg_spread};
  std::fwrite(s, 8, 5, o);
  std::fwrite(v.mean.data(), 8, (size_t)n, o);
  std::fwrite(v.spread.data(), 8, (size_t)n, o);
  std::fclose(o);
  return 0;
}
