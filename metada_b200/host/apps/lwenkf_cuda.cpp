// lwenkf <config>  on the CUDA backend (mirrors applications/data_assimilation/ensemble/lwenkf.cpp)
#include <fstream>

#include "LWEnKF.hpp"
#include "app_common.hpp"

int main(int argc, char** argv) {
  return runDriver("LWEnKF", argc, argv, [](auto& config, auto& ensemble, auto& obs, auto& obs_op) {
    fwk::LWEnKF<BackendTag> lw(ensemble, obs, obs_op, config);
    // optional reproducible perturbations: analysis.perturbation_file = raw float64 [obs][member]
    try {
      const std::string zf = config.GetSubsection("analysis").Get("perturbation_file").asString();
      std::ifstream f(zf, std::ios::binary);
      std::vector<double> Z(obs.size() * ensemble.Size());
      f.read(reinterpret_cast<char*>(Z.data()), static_cast<std::streamsize>(Z.size() * 8));
      if (f) lw.setObservationPerturbations(std::move(Z));
    } catch (...) {
    }
    lw.Analyse();
    lw.saveEnsemble();
    auto r = lw.getAnalysisResults();
    std::cout << "LWEnKF diagnostics: innovation_norm=" << r.innovation_norm << " background_spread=" << r.background_spread
              << " analysis_spread=" << r.analysis_spread << " max_gain=" << r.max_kalman_gain << " min_gain=" << r.min_kalman_gain
              << " cond=" << r.condition_number << " max_weight=" << r.max_weight << " min_weight=" << r.min_weight
              << " weight_variance=" << r.weight_variance << std::endl;
  });
}
