#pragma once
// Stochastic EnKF policy with the reference's interface (framework/algorithms/EnKF.hpp:84-134,139,
// 261,287): same AnalysisResults fields, same config keys (analysis.inflation, inflation_method,
// output_base_file, format).  Analyse() = mdc_enkf_analyse.  The reference draws its observation
// perturbations from an unseeded mt19937 (EnKF.hpp:346-347); here they come from a counter-based
// device generator seeded by the optional key analysis.seed (default 7) mixed with a per-call counter, or from
// setObservationPerturbations() for reproducible comparisons.
#include <string>
#include <vector>

#include "Config.hpp"
#include "DeviceAnalysis.hpp"
#include "Ensemble.hpp"
#include "Logger.hpp"
#include "ObsOperator.hpp"
#include "Observation.hpp"

namespace metada::framework {

template <typename BackendTag>
class EnKF {
 public:
  struct AnalysisResults {
    double innovation_norm;
    double analysis_increment_norm;
    double background_spread;
    double analysis_spread;
    double max_kalman_gain;
    double min_kalman_gain;
    double condition_number;
    int ensemble_size;
    int observation_count;
    std::string inflation_method;
    double inflation_factor;
  };

  EnKF(Ensemble<BackendTag>& ensemble, Observation<BackendTag>& obs,
       const ObsOperator<BackendTag>& obs_op, const Config<BackendTag>& config)
      : ensemble_(ensemble), obs_(obs), obs_op_(obs_op) {
    auto analysis_config = config.GetSubsection("analysis");
    inflation_factor_ = analysis_config.Get("inflation").asFloat();
    output_base_file_ = analysis_config.Get("output_base_file").asString();
    format_ = analysis_config.Get("format").asString();
    inflation_method_ = analysis_config.Get("inflation_method").asString();
    if (inflation_method_ != "multiplicative" && inflation_method_ != "additive" && inflation_method_ != "relaxation") {
      logger_.Warning() << "Unknown inflation method '" << inflation_method_ << "', using multiplicative inflation";
      inflation_method_ = "multiplicative";
    }
    try { seed_ = static_cast<uint64_t>(analysis_config.Get("seed").asInt()); } catch (...) {}
    try { resident_ = analysis_config.Get("resident").asBool(); } catch (...) {}   // DeviceAnalysis.hpp
    logger_.Info() << "EnKF constructed with " << ensemble_.Size() << " members (device path)";
  }

  /** Z: standard-normal draws, row-major [obs][member]; obs_pert = sqrt(R_ii) * Z (EnKF.hpp:340-361). */
  void setObservationPerturbations(std::vector<double> Z) { Z_ = std::move(Z); }

  void Analyse() {
    logger_.Info() << "EnKF analysis started";
    backends::cuda::DeviceObservations dobs(obs_.backend());
    if (!Z_.empty() && Z_.size() != dobs.size() * ensemble_.Size())
      throw std::invalid_argument("EnKF: observation perturbations must be [obs][member]");
    mdc_enkf_diag d{};
    device::analyseOnDevice(ensemble_, resident_, [&](backends::cuda::DeviceEnsemble& dev) {
      backends::cuda::DeviceContext::Instance().check(
          mdc_enkf_analyse(dev.get(), dobs.get(), inflation_factor_, Z_.empty() ? nullptr : Z_.data(),
                           seed_ + 0x9E3779B97F4A7C15ull * calls_++ /* a fresh stream every cycle */, 1, &d),
          "mdc_enkf_analyse");
    });                              // (the mean of EnKF.hpp:237 included)
    diag_ = d;
    logger_.Info() << "EnKF analysis completed";
  }

  void saveEnsemble() const {
    logger_.Info() << "EnKF saving ensemble";
    ensemble_.Mean().saveToFile(output_base_file_ + "_mean." + format_);
    for (size_t i = 0; i < ensemble_.Size(); ++i)
      ensemble_.GetMember(i).saveToFile(output_base_file_ + "_member_" + std::to_string(i) + "." + format_);
    logger_.Info() << "Diagnostics saved to: " << output_base_file_ + "_diagnostics.txt";
    logger_.Info() << "EnKF ensemble saved";
  }

  AnalysisResults getAnalysisResults() const {
    AnalysisResults r;
    r.innovation_norm = diag_.innovation_norm;
    r.analysis_increment_norm = 0.0;           // never assigned in the reference either (EnKF.hpp:398)
    r.background_spread = diag_.background_spread;
    r.analysis_spread = diag_.analysis_spread;
    r.max_kalman_gain = diag_.max_kalman_gain;
    r.min_kalman_gain = diag_.min_kalman_gain;
    r.condition_number = diag_.condition_number;
    r.ensemble_size = static_cast<int>(ensemble_.Size());
    r.observation_count = static_cast<int>(obs_.size());
    r.inflation_method = inflation_method_;
    r.inflation_factor = inflation_factor_;
    return r;
  }

 private:
  Ensemble<BackendTag>& ensemble_;
  Observation<BackendTag>& obs_;
  const ObsOperator<BackendTag>& obs_op_;
  std::string inflation_method_;
  double inflation_factor_ = 1.0;
  std::string output_base_file_;
  std::string format_ = "txt";
  uint64_t seed_ = 7, calls_ = 0;
  bool resident_ = true;
  std::vector<double> Z_;
  mdc_enkf_diag diag_{};
  Logger<BackendTag>& logger_ = Logger<BackendTag>::Instance();
};

}  // namespace metada::framework
