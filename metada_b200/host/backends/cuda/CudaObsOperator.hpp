#pragma once
// Observation-operator backend of the CUDA backend: H = 4-point inverse-distance interpolation,
// the same operator as backends/common/obsoperator/IdentityObsOperator.hpp:154-180, 594-676, run on
// the device (mdc_hx_idw4) and bit-identical to it.  Satisfies framework::ObsOperatorBackendImpl
// (ObsOperatorConcepts.hpp:35-60).  apply() keeps the reference's per-member signature
// (std::vector<double> of H(x) at every observation); the ensemble filters use the batched device
// path instead (algorithms/DeviceAnalysis.hpp) and never call it per grid point.
#include <string>
#include <vector>

#include "CudaApi.hpp"

namespace metada::backends::cuda {

template <typename StateBackend, typename ObsBackend, typename ControlVariableBackend>
class CudaObsOperator {
 public:
  CudaObsOperator() = delete;
  CudaObsOperator(const CudaObsOperator&) = delete;
  CudaObsOperator& operator=(const CudaObsOperator&) = delete;
  CudaObsOperator(CudaObsOperator&&) noexcept = default;
  CudaObsOperator& operator=(CudaObsOperator&&) noexcept = default;

  template <typename ConfigBackend>
  explicit CudaObsOperator(const ConfigBackend& config) { initialize(config); }
  template <typename ConfigBackend>
  CudaObsOperator(const ConfigBackend& config, [[maybe_unused]] const ControlVariableBackend& cb) { initialize(config); }

  template <typename ConfigBackend>
  void initialize(const ConfigBackend& config) {
    if (initialized_) throw std::runtime_error("CudaObsOperator already initialized");
    try { required_state_vars_ = config.Get("required_state_vars").asVectorString(); } catch (...) { required_state_vars_ = {"state"}; }
    try { required_obs_vars_ = config.Get("required_obs_vars").asVectorString(); } catch (...) { required_obs_vars_ = {}; }
    initialized_ = true;
  }
  bool isInitialized() const { return initialized_; }

  std::vector<double> apply(const StateBackend& state, const ObsBackend& obs) const {
    if (!initialized_) throw std::runtime_error("CudaObsOperator not initialized");
    const auto& g = state.geometry();
    DeviceEnsemble ens(g.x_dim(), g.y_dim(), g.z_dim(), 1);
    ens.upload({static_cast<const double*>(state.getData())});
    DeviceObservations dobs(obs);
    auto& c = DeviceContext::Instance();
    c.check(mdc_hx_idw4(ens.get(), dobs.get()), "mdc_hx_idw4");
    std::vector<double> y(dobs.size());
    c.check(mdc_hx_download(dobs.get(), y.data(), nullptr, nullptr, nullptr), "mdc_hx_download");
    return y;
  }

  const std::vector<std::string>& getRequiredStateVars() const { return required_state_vars_; }
  const std::vector<std::string>& getRequiredObsVars() const { return required_obs_vars_; }

 private:
  bool initialized_ = false;
  std::vector<std::string> required_state_vars_, required_obs_vars_;
};

}  // namespace metada::backends::cuda
