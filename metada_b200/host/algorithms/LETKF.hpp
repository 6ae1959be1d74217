#pragma once
// LETKF policy with the reference's interface (framework/algorithms/LETKF.hpp:38-57, 63, 125):
//   LETKF(Ensemble&, Observation&, const ObsOperator&, const Config&); Analyse(); saveEnsemble()
// Analyse() runs on the device: H(x) once from the background ensemble, bucketed local-observation
// index, one CTA per grid column (mdc_letkf_analyse).  Config keys are the reference's
// (inflation, localization_radius, output_base_file, format -- all via asFloat/asString, so reals
// arrive float-narrowed exactly as in LETKF.hpp:52-55) plus the optional
//   mode: "ref_compat" (default; the arithmetic of LETKF.hpp:209-238) | "ref_etkf" | "canonical"
//   localization_function: "cutoff" | "gaspari_cohn" | "gaussian" | "exponential" | "ref_gaspari_cohn"
//                          (canonical only; default gaspari_cohn; the last three are LWEnKF.hpp:597-635)
//   localization_scale:    length scale of the last three (default: localization_radius)
//   vertical_radius: levels (canonical only; default 0 = none)
#include <string>

#include "Config.hpp"
#include "DeviceAnalysis.hpp"
#include "Ensemble.hpp"
#include "Logger.hpp"
#include "ObsOperator.hpp"
#include "Observation.hpp"

namespace metada::framework {

template <typename BackendTag>
class LETKF {
 public:
  LETKF(Ensemble<BackendTag>& ensemble, Observation<BackendTag>& obs,
        const ObsOperator<BackendTag>& obs_op, const Config<BackendTag>& config)
      : ensemble_(ensemble),
        obs_(obs),
        obs_op_(obs_op),
        inflation_(config.Get("inflation").asFloat()),
        localization_radius_(config.Get("localization_radius").asFloat()),
        output_base_file_(config.Get("output_base_file").asString()),
        format_(config.Get("format").asString()) {
    params_ = mdc_letkf_params{};
    params_.radius = localization_radius_;
    params_.inflation = inflation_;
    params_.mode = MDC_MODE_REF_COMPAT;
    params_.loc = MDC_LOC_CUTOFF;
    params_.use_R = 1;
    std::string mode = "ref_compat", loc = "gaspari_cohn";
    try { mode = config.Get("mode").asString(); } catch (...) {}
    try { loc = config.Get("localization_function").asString(); } catch (...) {}
    try { params_.radius_v = config.Get("vertical_radius").asFloat(); } catch (...) {}
    try { params_.loc_scale = config.Get("localization_scale").asFloat(); } catch (...) {}
    if (mode == "ref_compat") params_.mode = MDC_MODE_REF_COMPAT;
    else if (mode == "ref_etkf") params_.mode = MDC_MODE_REF_ETKF;
    else if (mode == "canonical") {
      params_.mode = MDC_MODE_CANONICAL;
      if (loc == "cutoff") params_.loc = MDC_LOC_CUTOFF;
      else if (loc == "gaspari_cohn") params_.loc = MDC_LOC_GASPARI_COHN;
      else if (loc == "gaussian") params_.loc = MDC_LOC_GAUSSIAN;
      else if (loc == "exponential") params_.loc = MDC_LOC_EXPONENTIAL;
      else if (loc == "ref_gaspari_cohn") params_.loc = MDC_LOC_REF_GASPARI_COHN;
      else throw std::invalid_argument("LETKF: unknown localization_function '" + loc + "'");
    }
    else throw std::invalid_argument("LETKF: unknown mode '" + mode + "'");
    logger_.Info() << "LETKF constructed with radius " << localization_radius_ << " (device path, mode " << mode << ")";
  }

  void Analyse() {
    logger_.Info() << "LETKF analysis started";
    auto dev = device::uploadEnsemble(ensemble_);
    backends::cuda::DeviceObservations dobs(obs_.backend());
    auto& ctx = backends::cuda::DeviceContext::Instance();
    ctx.check(mdc_letkf_analyse(dev->get(), dobs.get(), &params_, &stats_), "mdc_letkf_analyse");
    device::downloadEnsemble(*dev, ensemble_);
    ensemble_.RecomputeMean();                       // LETKF.hpp:116
    logger_.Info() << "LETKF analysis completed: " << stats_.columns << " columns, mean local obs "
                   << (stats_.columns ? static_cast<double>(stats_.sum_local_obs) / stats_.columns : 0.0)
                   << ", device time " << stats_.ms_total << " ms";
  }

  void saveEnsemble() const {
    logger_.Info() << "LETKF saving ensemble";
    ensemble_.Mean().saveToFile(output_base_file_ + "_mean." + format_);
    for (size_t i = 0; i < ensemble_.Size(); ++i)
      ensemble_.GetMember(i).saveToFile(output_base_file_ + "_member_" + std::to_string(i) + "." + format_);
    logger_.Info() << "LETKF ensemble saved";
  }

  const mdc_letkf_stats& deviceStats() const { return stats_; }

 private:
  Ensemble<BackendTag>& ensemble_;
  Observation<BackendTag>& obs_;
  const ObsOperator<BackendTag>& obs_op_;
  double inflation_;
  double localization_radius_;
  std::string output_base_file_ = "analysis";
  std::string format_ = "nc";
  mdc_letkf_params params_{};
  mdc_letkf_stats stats_{};
  Logger<BackendTag>& logger_ = Logger<BackendTag>::Instance();
};

}  // namespace metada::framework
