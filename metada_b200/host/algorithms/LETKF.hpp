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
//   streaming: "auto" (default) | "on" | "off" -- "on" / "auto" with GRID observations: the members' host arrays
//              (State::getDataPtr) go through mdc_stream_analyse in row slabs, in place (an ensemble larger than
//              the device is fine); "off" or geographic observations: upload all, analyse, download all
//   slab_rows: rows per slab of the streamed path (default: chosen by the runtime)
//   resident:  true (default) | false -- one-shot path only: the ensemble stays in its device store across Analyse()
//              calls; members come back lazily when the host reads them and only members the host may have written
//              are uploaded again (DeviceAnalysis.hpp); false: upload all, analyse, download all, every call
// Several processes (one per GPU; column sharding with the NCCL observation halo): environment MDC_RANK,
// MDC_WORLD_SIZE, MDC_COMM_ID_FILE (rank 0 writes the NCCL id there, the others wait for it), MDC_DEVICE (default:
// rank).  Every process reads the whole ensemble, analyses its rows, and receives the others' rows afterwards
// (mdc_comm_allgather_rows), so saveEnsemble() writes the same files on every rank (give them different
// output_base_file, or save on rank 0 only).
#include <chrono>
#include <cstdlib>
#include <fstream>
#include <string>
#include <thread>

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
    try { streaming_ = config.Get("streaming").asString(); } catch (...) {}
    try { slab_rows_ = config.Get("slab_rows").asInt(); } catch (...) {}
    try { resident_ = config.Get("resident").asBool(); } catch (...) {}
    if (streaming_ != "auto" && streaming_ != "on" && streaming_ != "off")
      throw std::invalid_argument("LETKF: streaming must be auto, on or off");
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
    if (streaming_ != "off" && analyseStreamed()) {
      ensemble_.RecomputeMean();                     // LETKF.hpp:116
    } else {
      backends::cuda::DeviceObservations dobs(obs_.backend());
      device::analyseOnDevice(ensemble_, resident_, [&](backends::cuda::DeviceEnsemble& dev) {
        backends::cuda::DeviceContext::Instance().check(mdc_letkf_analyse(dev.get(), dobs.get(), &params_, &stats_),
                                                        "mdc_letkf_analyse");
      });                                            // (the mean of LETKF.hpp:116 included)
    }
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
  // Streamed (and, with MDC_WORLD_SIZE > 1, sharded) analysis straight from the members' host arrays.  false: not
  // applicable (geographic observations / geography, or a plain "auto" run that fits one shot) -- the caller falls back.
  bool analyseStreamed() {
    const auto* geometry = ensemble_.GetMember(0).geometry();
    if (!geometry) throw std::runtime_error("Geometry pointer is null in device analysis");
    const auto& g = geometry->backend();
    std::vector<int32_t> x, y, z;
    std::vector<double> val, err;
    std::vector<uint8_t> valid;
    for (const auto& p : obs_.backend()) {
      if (p.location.getCoordinateSystem() != framework::CoordinateSystem::GRID) return false;   // (haversine path: one shot)
      auto [i, j, l] = p.location.getGridCoords();
      x.push_back(i); y.push_back(j); z.push_back(l);
      val.push_back(p.value); err.push_back(p.error); valid.push_back(p.is_valid ? 1 : 0);
    }
    const backends::cuda::ProcessGroup pg;
    const int world = pg.world, rank = pg.rank;
    const int gnx = static_cast<int>(g.x_dim()), gny = static_cast<int>(g.y_dim()), nz = static_cast<int>(g.z_dim());
    const int k = static_cast<int>(ensemble_.Size());
    const double bytes = 8.0 * gnx * gny * nz * k;
    if (streaming_ == "auto" && world == 1 && bytes < 2e9) return false;      // small: the one-shot path has less overhead
    mdc_stream_config cfg{};
    cfg.gnx = gnx; cfg.gny = gny; cfg.nz = nz; cfg.k = k;
    cfg.row0 = static_cast<int>((static_cast<long long>(gny) * rank) / world);
    cfg.row1 = static_cast<int>((static_cast<long long>(gny) * (rank + 1)) / world);
    cfg.slab_rows = slab_rows_; cfg.slots = 4; cfg.sm_reserve = 8; cfg.radius = params_.radius;
    const int device = pg.device;
    mdc_stream* st = nullptr;
    if (mdc_stream_create(device, &cfg, &st)) throw std::runtime_error("mdc_stream_create failed");
    auto fail = [&](const char* what) {
      std::string msg = std::string(what) + ": " + mdc_stream_last_error(st);
      mdc_stream_destroy(st);
      throw std::runtime_error(msg);
    };
    try { pg.attach(st); } catch (const std::exception& e) { fail(e.what()); }
    std::vector<double*> ptrs;
    for (int m = 0; m < k; ++m) ptrs.push_back(ensemble_.GetMember(m).template getDataPtr<double>());
    if (mdc_stream_analyse(st, ptrs.data(), 0, gny, static_cast<int64_t>(val.size()), x.data(), y.data(), z.data(), val.data(),
                           err.data(), valid.data(), &params_, &stats_))
      fail("mdc_stream_analyse");
    if (world > 1 && mdc_comm_allgather_rows(st, ptrs.data())) fail("mdc_comm_allgather_rows");
    logger_.Info() << "LETKF streamed analysis: " << mdc_stream_slabs(st) << " slabs, rank " << rank << " of " << world;
    mdc_stream_destroy(st);
    return true;
  }

  Ensemble<BackendTag>& ensemble_;
  Observation<BackendTag>& obs_;
  const ObsOperator<BackendTag>& obs_op_;
  double inflation_;
  double localization_radius_;
  std::string output_base_file_ = "analysis";
  std::string format_ = "nc";
  mdc_letkf_params params_{};
  mdc_letkf_stats stats_{};
  std::string streaming_ = "auto";
  int slab_rows_ = 0;
  bool resident_ = true;
  Logger<BackendTag>& logger_ = Logger<BackendTag>::Instance();
};

}  // namespace metada::framework
