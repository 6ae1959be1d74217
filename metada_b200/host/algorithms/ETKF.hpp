#pragma once
// Global ETKF policy with the reference's interface (framework/algorithms/ETKF.hpp:86-95,100,185);
// Analyse() = mdc_etkf_analyse (ensemble-space solve on one CTA + streaming X'W update).
#include <string>

#include "Config.hpp"
#include "DeviceAnalysis.hpp"
#include "Ensemble.hpp"
#include "Logger.hpp"
#include "ObsOperator.hpp"
#include "Observation.hpp"

namespace metada::framework {

template <typename BackendTag>
class ETKF {
 public:
  ETKF(Ensemble<BackendTag>& ensemble, Observation<BackendTag>& obs,
       const ObsOperator<BackendTag>& obs_op, const Config<BackendTag>& config)
      : ensemble_(ensemble), obs_(obs), obs_op_(obs_op),
        inflation_(config.Get("inflation").asFloat()),
        output_base_file_(config.Get("output_base_file").asString()),
        format_(config.Get("format").asString()) {
    try { resident_ = config.Get("resident").asBool(); } catch (...) {}   // DeviceAnalysis.hpp
    logger_.Info() << "ETKF constructed (device path)";
  }

  void Analyse() {
    logger_.Info() << "ETKF analysis started";
    backends::cuda::DeviceObservations dobs(obs_.backend());
    // (ends with the ensemble mean, so that saveEnsemble() can use Mean(): ETKF.hpp:190)
    device::analyseOnDevice(ensemble_, resident_, [&](backends::cuda::DeviceEnsemble& dev) {
      backends::cuda::DeviceContext::Instance().check(mdc_etkf_analyse(dev.get(), dobs.get(), inflation_), "mdc_etkf_analyse");
    });
    logger_.Info() << "ETKF analysis completed";
  }

  void saveEnsemble() const {
    logger_.Info() << "ETKF saving ensemble";
    ensemble_.Mean().saveToFile(output_base_file_ + "_mean." + format_);
    for (size_t i = 0; i < ensemble_.Size(); ++i)
      ensemble_.GetMember(i).saveToFile(output_base_file_ + "_member_" + std::to_string(i) + "." + format_);
    logger_.Info() << "ETKF ensemble saved";
  }

 private:
  Ensemble<BackendTag>& ensemble_;
  Observation<BackendTag>& obs_;
  const ObsOperator<BackendTag>& obs_op_;
  double inflation_;
  std::string output_base_file_;
  std::string format_ = "txt";
  bool resident_ = true;
  Logger<BackendTag>& logger_ = Logger<BackendTag>::Instance();
};

}  // namespace metada::framework
