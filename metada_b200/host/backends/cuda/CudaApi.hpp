#pragma once
// RAII + exception layer over the C ABI (include/metada_cuda_c_api.h), following the reference's
// convention for native bridges: opaque handle in a smart pointer, non-zero return code -> throw
// std::runtime_error (backends/wrf/WRFObsOperator.hpp:150-154 pattern).
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "metada_cuda_c_api.h"

namespace metada::backends::cuda {

class DeviceContext {
 public:
  /** One context per process/device; created on first use. No CPU fallback: throws without a GPU. */
  static DeviceContext& Instance(int device = 0) {
    static DeviceContext ctx(device);
    return ctx;
  }
  mdc_ctx* get() const { return ctx_; }
  void check(int rc, const char* what) const {
    if (rc != MDC_OK)
      throw std::runtime_error(std::string(what) + " failed (rc=" + std::to_string(rc) + "): " +
                               mdc_last_error(ctx_));
  }
  ~DeviceContext() { mdc_ctx_destroy(ctx_); }
  DeviceContext(const DeviceContext&) = delete;
  DeviceContext& operator=(const DeviceContext&) = delete;

 private:
  explicit DeviceContext(int device) {
    if (mdc_ctx_create(device, &ctx_) != MDC_OK || !ctx_)
      throw std::runtime_error("CUDA backend: no usable CUDA device (this backend has no CPU fallback)");
  }
  mdc_ctx* ctx_ = nullptr;
};

/** Device-resident ensemble store [col][lev][member] (replaces vector<State> for the analysis). */
class DeviceEnsemble {
 public:
  DeviceEnsemble(int nx, int ny, int nz, int k) : nx_(nx), ny_(ny), nz_(nz), k_(k) {
    auto& c = DeviceContext::Instance();
    c.check(mdc_ens_create(c.get(), nx, ny, nz, k, &h_), "mdc_ens_create");
  }
  ~DeviceEnsemble() { mdc_ens_destroy(h_); }
  DeviceEnsemble(const DeviceEnsemble&) = delete;
  DeviceEnsemble& operator=(const DeviceEnsemble&) = delete;
  mdc_ens* get() const { return h_; }
  int members() const { return k_; }
  size_t pointsPerMember() const { return static_cast<size_t>(nx_) * ny_ * nz_; }
  void upload(const std::vector<const double*>& members) {
    DeviceContext::Instance().check(mdc_ens_upload_members(h_, 0, static_cast<int>(members.size()), members.data()),
                                    "mdc_ens_upload_members");
  }
  void download(const std::vector<double*>& members) {
    DeviceContext::Instance().check(mdc_ens_download_members(h_, 0, static_cast<int>(members.size()), members.data()),
                                    "mdc_ens_download_members");
  }
  void mean(double* host) { DeviceContext::Instance().check(mdc_ens_mean(h_, host), "mdc_ens_mean"); }

 private:
  mdc_ens* h_ = nullptr;
  int nx_, ny_, nz_, k_;
};

/** Device SoA mirror of an observation backend (anything iterable yielding
 *  {location, value, error, is_valid} like GridObservation / PointObservation.hpp:63-67). */
class DeviceObservations {
 public:
  template <typename ObsBackend>
  explicit DeviceObservations(const ObsBackend& obs) {
    std::vector<int32_t> x, y, z;
    std::vector<double> val, err;
    std::vector<uint8_t> valid;
    for (const auto& p : obs) {
      auto [i, j, k] = p.location.getGridCoords();   // throws for non-GRID locations, like distance_to
      x.push_back(i); y.push_back(j); z.push_back(k);
      val.push_back(p.value); err.push_back(p.error); valid.push_back(p.is_valid ? 1 : 0);
    }
    size_ = x.size();
    auto& c = DeviceContext::Instance();
    c.check(mdc_obs_create(c.get(), static_cast<int64_t>(size_), x.data(), y.data(), z.data(), val.data(),
                           err.data(), valid.data(), nullptr, &h_),
            "mdc_obs_create");
  }
  ~DeviceObservations() { mdc_obs_destroy(h_); }
  DeviceObservations(const DeviceObservations&) = delete;
  DeviceObservations& operator=(const DeviceObservations&) = delete;
  mdc_obs* get() const { return h_; }
  size_t size() const { return size_; }

 private:
  mdc_obs* h_ = nullptr;
  size_t size_ = 0;
};

}  // namespace metada::backends::cuda
