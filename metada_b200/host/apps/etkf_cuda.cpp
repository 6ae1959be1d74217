// etkf <config>  on the CUDA backend (mirrors applications/data_assimilation/ensemble/etkf.cpp:45-62)
#include "ETKF.hpp"
#include "app_common.hpp"

int main(int argc, char** argv) {
  return runDriver("ETKF", argc, argv, [](auto& config, auto& ensemble, auto& obs, auto& obs_op) {
    fwk::ETKF<BackendTag> etkf(ensemble, obs, obs_op, config.GetSubsection("analysis"));
    etkf.Analyse();
    etkf.saveEnsemble();
  });
}
