// enkf <config>  on the CUDA backend (mirrors applications/data_assimilation/ensemble/enkf.cpp:57-88)
#include <fstream>

#include "EnKF.hpp"
#include "app_common.hpp"

int main(int argc, char** argv) {
  return runDriver("EnKF", argc, argv, [](auto& config, auto& ensemble, auto& obs, auto& obs_op) {
    fwk::EnKF<BackendTag> enkf(ensemble, obs, obs_op, config);
    // optional reproducible perturbations: analysis.perturbation_file = raw float64 [obs][member]
    try {
      const std::string zf = config.GetSubsection("analysis").Get("perturbation_file").asString();
      std::ifstream f(zf, std::ios::binary);
      std::vector<double> Z(obs.size() * ensemble.Size());
      f.read(reinterpret_cast<char*>(Z.data()), static_cast<std::streamsize>(Z.size() * 8));
      if (f) enkf.setObservationPerturbations(std::move(Z));
    } catch (...) {
    }
    enkf.Analyse();
    enkf.saveEnsemble();
    auto r = enkf.getAnalysisResults();
    std::cout << "EnKF diagnostics: innovation_norm=" << r.innovation_norm << " background_spread=" << r.background_spread
              << " analysis_spread=" << r.analysis_spread << " max_gain=" << r.max_kalman_gain
              << " min_gain=" << r.min_kalman_gain << " cond=" << r.condition_number << std::endl;
  });
}
