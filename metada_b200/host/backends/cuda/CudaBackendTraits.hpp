#pragma once
// BackendTraits<CudaBackendTag>: wires the CUDA backend into METADA's trait-injection scheme
// exactly like traits/SimpleBackendTraits.hpp:53-105 does for the Simple backend.  Observation
// storage, config and logger backends are the reference's own generic ones.
#include "BackendTraits.hpp"
#include "GridObservation.hpp"
#include "GridObservationIterator.hpp"

#ifdef CONFIG_BACKEND_JSON
#include "JsonConfig.hpp"
#else
#include "YamlConfig.hpp"
#endif
#ifdef LOGGER_BACKEND_CONSOLE
#include "ConsoleLogger.hpp"
#else
#include "NgLogger.hpp"
#endif

#include "CudaGeometry.hpp"
#include "CudaObsOperator.hpp"
#include "CudaState.hpp"

namespace metada::traits {

struct CudaBackendTag {};

template <>
struct BackendTraits<CudaBackendTag> {
#ifdef CONFIG_BACKEND_JSON
  using ConfigBackend = backends::config::JsonConfig;
#else
  using ConfigBackend = backends::config::YamlConfig;
#endif
#ifdef LOGGER_BACKEND_CONSOLE
  using LoggerBackend = backends::logger::ConsoleLogger<ConfigBackend>;
#else
  using LoggerBackend = backends::logger::NgLogger<ConfigBackend>;
#endif
  using GeometryBackend = backends::cuda::CudaGeometry;
  using GeometryIteratorBackend = backends::cuda::CudaGeometryIterator;
  using StateBackend = backends::cuda::CudaState;
  using IncrementBackend = backends::cuda::CudaIncrement;
  using ControlVariableBackend = backends::cuda::CudaIncrement;
  using ObservationBackend = backends::common::observation::GridObservation;
  using ObservationIteratorBackend = backends::common::observation::GridObservationIterator;
  using ObsOperatorBackend =
      backends::cuda::CudaObsOperator<StateBackend, ObservationBackend, ControlVariableBackend>;
};

}  // namespace metada::traits
