#pragma once
// Locally weighted EnKF policy with the reference's interface (framework/algorithms/LWEnKF.hpp:103-126 AnalysisResults,
// :134-202 constructor and config keys -- analysis.inflation, localization_radius, inflation_method,
// localization_function (gaussian | exponential | cutoff | gaspari_cohn), weighting_scheme (uniform | adaptive |
// inverse_var | likelihood), output_base_file, format --, :207 Analyse, :339 saveEnsemble, :365 getAnalysisResults).
// Analyse() = mdc_lwenkf_analyse.  The reference draws its observation perturbations from an unseeded mt19937
// (:670-672); here they come from the device generator seeded by analysis.seed (default 7) mixed with a per-call
// counter, or from setObservationPerturbations() for reproducible comparisons.
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
class LWEnKF {
 public:
  struct AnalysisResults {
    double innovation_norm;
    double analysis_increment_norm;
    double background_spread;
    double analysis_spread;
    double max_kalman_gain;
    double min_kalman_gain;
    double condition_number;
    double localization_radius;
    std::string localization_function;
    std::string weighting_scheme;
    double max_weight;
    double min_weight;
    double weight_variance;
    int ensemble_size;
    int observation_count;
    std::string inflation_method;
    double inflation_factor;
  };

  LWEnKF(Ensemble<BackendTag>& ensemble, Observation<BackendTag>& obs, const ObsOperator<BackendTag>& obs_op,
         const Config<BackendTag>& config)
      : ensemble_(ensemble), obs_(obs), obs_op_(obs_op) {
    auto analysis_config = config.GetSubsection("analysis");
    inflation_factor_ = analysis_config.Get("inflation").asFloat();
    localization_radius_ = analysis_config.Get("localization_radius").asFloat();
    output_base_file_ = analysis_config.Get("output_base_file").asString();
    format_ = analysis_config.Get("format").asString();
    inflation_method_ = analysis_config.Get("inflation_method").asString();
    if (inflation_method_ != "multiplicative" && inflation_method_ != "additive" && inflation_method_ != "relaxation") {
      logger_.Warning() << "Unknown inflation method '" << inflation_method_ << "', using multiplicative inflation";
      inflation_method_ = "multiplicative";
    }
    loc_function_ = analysis_config.Get("localization_function").asString();
    if (loc_function_ == "gaussian") loc_fn_ = MDC_LOC_GAUSSIAN;
    else if (loc_function_ == "exponential") loc_fn_ = MDC_LOC_EXPONENTIAL;
    else if (loc_function_ == "cutoff") loc_fn_ = MDC_LOC_CUTOFF;
    else if (loc_function_ == "gaspari_cohn") loc_fn_ = MDC_LOC_REF_GASPARI_COHN;
    else {
      logger_.Warning() << "Unknown localization function '" << loc_function_ << "', using Gaussian localization";
      loc_function_ = "gaussian";
      loc_fn_ = MDC_LOC_GAUSSIAN;
    }
    weighting_ = analysis_config.Get("weighting_scheme").asString();
    if (weighting_ == "uniform") weighting_code_ = MDC_LW_UNIFORM;
    else if (weighting_ == "adaptive") weighting_code_ = MDC_LW_ADAPTIVE;
    else if (weighting_ == "inverse_var") weighting_code_ = MDC_LW_INVERSE_VAR;
    else if (weighting_ == "likelihood") weighting_code_ = MDC_LW_LIKELIHOOD;
    else {
      logger_.Warning() << "Unknown weighting scheme '" << weighting_ << "', using uniform weighting";
      weighting_ = "uniform";
      weighting_code_ = MDC_LW_UNIFORM;
    }
    try { seed_ = static_cast<uint64_t>(analysis_config.Get("seed").asInt()); } catch (...) {}
    try { resident_ = analysis_config.Get("resident").asBool(); } catch (...) {}   // DeviceAnalysis.hpp
    logger_.Info() << "LWEnKF constructed with " << ensemble_.Size() << " members (device path)";
  }

  /** Z: standard-normal draws, row-major [obs][member]; obs_pert = sqrt(R_ii) * Z (LWEnKF.hpp:665-685). */
  void setObservationPerturbations(std::vector<double> Z) { Z_ = std::move(Z); }

  void Analyse() {
    logger_.Info() << "LWEnKF analysis started";
    backends::cuda::DeviceObservations dobs(obs_.backend());
    if (!Z_.empty() && Z_.size() != dobs.size() * ensemble_.Size())
      throw std::invalid_argument("LWEnKF: observation perturbations must be [obs][member]");
    mdc_lwenkf_diag d{};
    const uint64_t seed = seed_ + 0x9E3779B97F4A7C15ull * calls_++;      // a fresh stream every cycle
    device::analyseOnDevice(ensemble_, resident_, [&](backends::cuda::DeviceEnsemble& dev) {
      backends::cuda::DeviceContext::Instance().check(
          mdc_lwenkf_analyse(dev.get(), dobs.get(), inflation_factor_, localization_radius_, loc_fn_, weighting_code_,
                             Z_.empty() ? nullptr : Z_.data(), seed, &d),
          "mdc_lwenkf_analyse");
    });                              // (the mean of LWEnKF.hpp:317 included)
    diag_ = d;
    logger_.Info() << "LWEnKF analysis completed";
  }

  void saveEnsemble() const {
    logger_.Info() << "LWEnKF saving ensemble";
    ensemble_.Mean().saveToFile(output_base_file_ + "_mean." + format_);
    for (size_t i = 0; i < ensemble_.Size(); ++i)
      ensemble_.GetMember(i).saveToFile(output_base_file_ + "_member_" + std::to_string(i) + "." + format_);
    logger_.Info() << "LWEnKF ensemble saved";
  }

  AnalysisResults getAnalysisResults() const {
    AnalysisResults r;
    r.innovation_norm = diag_.innovation_norm;
    r.analysis_increment_norm = 0.0;           // never assigned in the reference either
    r.background_spread = diag_.background_spread;
    r.analysis_spread = diag_.analysis_spread;
    r.max_kalman_gain = diag_.max_kalman_gain;
    r.min_kalman_gain = diag_.min_kalman_gain;
    r.condition_number = diag_.condition_number;
    r.localization_radius = localization_radius_;
    r.localization_function = loc_function_;
    r.weighting_scheme = weighting_;
    r.max_weight = diag_.max_weight;
    r.min_weight = diag_.min_weight;
    r.weight_variance = diag_.weight_variance;
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
  std::string inflation_method_, loc_function_, weighting_;
  double inflation_factor_ = 1.0, localization_radius_ = 1.0;
  int loc_fn_ = MDC_LOC_GAUSSIAN, weighting_code_ = MDC_LW_UNIFORM;
  std::string output_base_file_;
  std::string format_ = "txt";
  uint64_t seed_ = 7, calls_ = 0;
  bool resident_ = true;
  std::vector<double> Z_;
  mdc_lwenkf_diag diag_{};
  Logger<BackendTag>& logger_ = Logger<BackendTag>::Instance();
};

}  // namespace metada::framework
