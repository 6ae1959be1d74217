#pragma once
// Shared body of the three ensemble drivers bound to CudaBackendTag.  Construction order and
// config sections follow applications/data_assimilation/ensemble/{letkf,etkf,enkf}.cpp:
// geometry -> ensemble -> observations -> identity control backend -> obs_operator -> algorithm.
#include <cstdint>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>

#include "ApplicationContext.hpp"
#include "Config.hpp"
#include "ControlVariableBackend.hpp"
#include "CudaBackendTraits.hpp"
#include "Ensemble.hpp"
#include "Geometry.hpp"
#include "IdentityControlVariableBackend.hpp"
#include "ObsOperator.hpp"
#include "Observation.hpp"

namespace fwk = metada::framework;
using BackendTag = metada::traits::CudaBackendTag;

// optional full-precision dump (the backend's saveToFile keeps the reference's 6-decimal text)
inline void dumpEnsemble(fwk::Ensemble<BackendTag>& ens, const std::string& path) {
  std::ofstream f(path, std::ios::binary);
  const int64_t k = static_cast<int64_t>(ens.Size()), n = static_cast<int64_t>(ens.GetMember(0).size());
  f.write(reinterpret_cast<const char*>(&k), 8);
  f.write(reinterpret_cast<const char*>(&n), 8);
  for (int64_t m = 0; m < k; ++m)
    f.write(reinterpret_cast<const char*>(ens.GetMember(m).template getDataPtr<double>()), n * 8);
}

template <typename MakeAndRun>
int runDriver(const char* name, int argc, char** argv, MakeAndRun&& body) {
  try {
    if (argc != 2 && argc != 4) {
      std::cerr << "Usage: " << name << " <config_file> [--dump <binary_file>]" << std::endl;
      return 1;
    }
    auto context = fwk::ApplicationContext<BackendTag>(2, argv);
    auto& logger = context.getLogger();
    auto& config = context.getConfig();
    logger.Info() << name << " application starting...";
    fwk::Geometry<BackendTag> geometry(config.GetSubsection("geometry"));
    fwk::Ensemble<BackendTag> ensemble(config.GetSubsection("ensemble"), geometry);
    fwk::Observation<BackendTag> observations(config.GetSubsection("observations"));
    auto control_backend = std::make_shared<fwk::IdentityControlVariableBackend<BackendTag>>();
    fwk::ObsOperator<BackendTag> obs_operator(config.GetSubsection("obs_operator"), *control_backend);
    body(config, ensemble, observations, obs_operator);
    if (argc == 4 && std::string(argv[2]) == "--dump") dumpEnsemble(ensemble, argv[3]);
    logger.Info() << name << " application completed successfully";
    return 0;
  } catch (const std::exception& e) {
    std::cerr << name << " application failed: " << e.what() << std::endl;
    return 1;
  }
}
