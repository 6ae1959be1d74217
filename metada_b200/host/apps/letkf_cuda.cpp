// letkf <config>  on the CUDA backend (mirrors applications/data_assimilation/ensemble/letkf.cpp:30-72)
#include "LETKF.hpp"
#include "app_common.hpp"

int main(int argc, char** argv) {
  return runDriver("LETKF", argc, argv, [](auto& config, auto& ensemble, auto& obs, auto& obs_op) {
    fwk::LETKF<BackendTag> letkf(ensemble, obs, obs_op, config.GetSubsection("analysis"));
    letkf.Analyse();
    letkf.saveEnsemble();
  });
}
