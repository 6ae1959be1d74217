// Two LETKF analyses with a "forecast" in between that touches ONE member on the host: shows (and lets the test
// check) what device residency moves over PCIe -- k members up for the first analysis, one member down + up for the
// forecast, nothing else until somebody reads the analysis on the host.  `analysis.resident: false` runs the same
// cycle through upload-all / download-all; the dumps must be byte-identical.
#include <cstdio>

#include "LETKF.hpp"
#include "app_common.hpp"

int main(int argc, char** argv) {
  return runDriver("LETKF resident cycle", argc, argv, [](auto& config, auto& ensemble, auto& obs, auto& obs_op) {
    fwk::LETKF<BackendTag> letkf(ensemble, obs, obs_op, config.GetSubsection("analysis"));
    auto counters = [&](const char* when) {
      const auto& link = ensemble.GetMember(0).backend().resident();
      if (link) std::printf("RESIDENT %s up %zu down %zu\n", when, link->store->membersUploaded(), link->store->membersDownloaded());
      else std::printf("RESIDENT %s up -1 down -1\n", when);
    };
    letkf.Analyse();
    counters("analysis1");
    // the "forecast": member 1 is advanced on the host (read-modify-write through the adapter's pointer), member 0
    // is only looked at
    double* x = ensemble.GetMember(1).template getDataPtr<double>();
    for (size_t i = 0; i < ensemble.GetMember(1).size(); ++i) x[i] += 0.125;
    const double peek = std::as_const(ensemble).GetMember(0).template getDataPtr<double>()[0];
    std::printf("RESIDENT peek %.17g\n", peek);
    counters("forecast");
    letkf.Analyse();
    counters("analysis2");
    std::printf("RESIDENT mean0 %.17g\n", std::as_const(ensemble).Mean().template getDataPtr<double>()[0]);
    letkf.saveEnsemble();
    counters("saved");
  });
}
