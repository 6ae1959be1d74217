#pragma once
// RAII + exception layer over the C ABI (include/metada_cuda_c_api.h), following the reference's
// convention for native bridges: opaque handle in a smart pointer, non-zero return code -> throw
// std::runtime_error (backends/wrf/WRFObsOperator.hpp:150-154 pattern).
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "Location.hpp"
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

/** Several processes, one per GPU (MDC_RANK, MDC_WORLD_SIZE, MDC_COMM_ID_FILE, MDC_DEVICE): rank 0 writes the NCCL id
 *  to the file, the others wait for it; attaches the communicator to a streaming handle created for the rank's rows. */
struct ProcessGroup {
  int rank = 0, world = 1, device = 0;
  ProcessGroup() {
    if (const char* w = std::getenv("MDC_WORLD_SIZE")) world = std::atoi(w);
    if (const char* r = std::getenv("MDC_RANK")) rank = std::atoi(r);
    device = std::getenv("MDC_DEVICE") ? std::atoi(std::getenv("MDC_DEVICE")) : (world > 1 ? rank : 0);
  }
  int row0(int gny) const { return static_cast<int>((static_cast<long long>(gny) * rank) / world); }
  int row1(int gny) const { return static_cast<int>((static_cast<long long>(gny) * (rank + 1)) / world); }
  /** throws std::runtime_error with the handle's message on failure (the caller destroys the handle) */
  void attach(mdc_stream* st) const {
    if (world < 2) return;
    const char* idfile = std::getenv("MDC_COMM_ID_FILE");
    if (!idfile) throw std::runtime_error("MDC_WORLD_SIZE > 1 needs MDC_COMM_ID_FILE");
    char id[128];
    if (rank == 0) {
      if (mdc_comm_get_unique_id(id, 128)) throw std::runtime_error("mdc_comm_get_unique_id failed");
      std::ofstream(std::string(idfile) + ".tmp", std::ios::binary).write(id, 128);
      std::rename((std::string(idfile) + ".tmp").c_str(), idfile);
    } else {
      for (int tries = 0;; ++tries) {
        std::ifstream f(idfile, std::ios::binary);
        if (f.read(id, 128)) break;
        if (tries > 6000) throw std::runtime_error("timed out waiting for MDC_COMM_ID_FILE");
        std::this_thread::sleep_for(std::chrono::milliseconds(10));
      }
    }
    if (mdc_comm_init(st, id, rank, world)) throw std::runtime_error(std::string("mdc_comm_init: ") + mdc_stream_last_error(st));
  }
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
    members_uploaded_ += members.size();
  }
  void download(const std::vector<double*>& members) {
    DeviceContext::Instance().check(mdc_ens_download_members(h_, 0, static_cast<int>(members.size()), members.data()),
                                    "mdc_ens_download_members");
    members_downloaded_ += members.size();
  }
  /** One member (run of members) at a time: what the lazily synchronised host views use. */
  void uploadMembers(int m0, const std::vector<const double*>& members) {
    DeviceContext::Instance().check(mdc_ens_upload_members(h_, m0, static_cast<int>(members.size()), members.data()),
                                    "mdc_ens_upload_members");
    members_uploaded_ += members.size();
  }
  void downloadMember(int m, double* host) {
    double* p[1] = {host};
    DeviceContext::Instance().check(mdc_ens_download_members(h_, m, 1, p), "mdc_ens_download_members");
    members_downloaded_ += 1;
  }
  void mean(double* host) { DeviceContext::Instance().check(mdc_ens_mean(h_, host), "mdc_ens_mean"); }
  int nx() const { return nx_; }
  int ny() const { return ny_; }
  int nz() const { return nz_; }
  /** Members moved over PCIe since construction (tests and logs: what residency saves). */
  size_t membersUploaded() const { return members_uploaded_; }
  size_t membersDownloaded() const { return members_downloaded_; }
  /** Column coordinates in degrees ([ny][nx], e.g. WRFGeometry::unstaggered_info().latitude_2d / longitude_2d)
   *  and the geometry's vertical coordinate: needed by observations with GEOGRAPHIC locations. */
  void setGeography(const std::vector<double>& lat, const std::vector<double>& lon, const std::vector<double>& vertical) {
    if (lat.size() != static_cast<size_t>(nx_) * ny_ || lon.size() != lat.size())
      throw std::invalid_argument("DeviceEnsemble::setGeography: coordinate arrays do not match the grid");
    DeviceContext::Instance().check(
        mdc_ens_set_geography(h_, lat.data(), lon.data(), static_cast<int>(vertical.size()), vertical.empty() ? nullptr : vertical.data()),
        "mdc_ens_set_geography");
  }
  /** Levels of each state variable of a [var][lev][y][x] member. */
  void setVariables(const std::vector<int32_t>& var_nlev) {
    DeviceContext::Instance().check(mdc_ens_set_variables(h_, static_cast<int>(var_nlev.size()), var_nlev.data()), "mdc_ens_set_variables");
  }

 private:
  mdc_ens* h_ = nullptr;
  int nx_, ny_, nz_, k_;
  size_t members_uploaded_ = 0, members_downloaded_ = 0;
};

/** Link between the host view of one member (CudaState) and its slot in a device-resident store
 *  (SURVEY "layout clash": what getDataPtr returns for a device-resident state = a host shadow, lazily
 *  synchronised).  After an analysis the device holds the truth (host_stale); the first host access of
 *  a member downloads that member only; a host access that may write marks the device copy stale, and
 *  the next analysis uploads just those members.  The store lives as long as any member refers to it. */
struct ResidentLink {
  std::shared_ptr<DeviceEnsemble> store;
  int member = -1;
  bool host_stale = false;
  bool device_stale = true;
};

/** Device SoA mirror of an observation backend (anything iterable yielding
 *  {location, value, error, is_valid} like GridObservation / PointObservation.hpp:63-67). */
class DeviceObservations {
 public:
  template <typename ObsBackend>
  explicit DeviceObservations(const ObsBackend& obs) {
    std::vector<int32_t> x, y, z;
    std::vector<double> lat, lon, lev, val, err;
    std::vector<uint8_t> valid;
    // GEOGRAPHIC locations (Location.hpp:82-84) select by haversine kilometres and are located on the grid by the
    // device (IdentityObsOperator.hpp:241-248); GRID locations are used as they are.  A mix throws, as
    // Location::distance_to does (Location.hpp:226-229).
    bool geographic = false, first = true;
    for (const auto& p : obs) {
      const bool g = p.location.getCoordinateSystem() == framework::CoordinateSystem::GEOGRAPHIC;
      if (first) { geographic = g; first = false; }
      if (g != geographic) throw std::runtime_error("DeviceObservations: GRID and GEOGRAPHIC locations cannot be mixed");
      if (g) {
        auto [la, lo, le] = p.location.getGeographicCoords();
        lat.push_back(la); lon.push_back(lo); lev.push_back(le);
      } else {
        auto [i, j, k] = p.location.getGridCoords();   // throws for other systems, like distance_to
        x.push_back(i); y.push_back(j); z.push_back(k);
      }
      val.push_back(p.value); err.push_back(p.error); valid.push_back(p.is_valid ? 1 : 0);
    }
    size_ = val.size();
    geographic_ = geographic;
    auto& c = DeviceContext::Instance();
    if (geographic)
      c.check(mdc_obs_create_geographic(c.get(), static_cast<int64_t>(size_), lat.data(), lon.data(), lev.data(), val.data(),
                                        err.data(), valid.data(), nullptr, &h_),
              "mdc_obs_create_geographic");
    else
      c.check(mdc_obs_create(c.get(), static_cast<int64_t>(size_), x.data(), y.data(), z.data(), val.data(),
                             err.data(), valid.data(), nullptr, &h_),
              "mdc_obs_create");
  }
  ~DeviceObservations() { mdc_obs_destroy(h_); }
  DeviceObservations(const DeviceObservations&) = delete;
  DeviceObservations& operator=(const DeviceObservations&) = delete;
  mdc_obs* get() const { return h_; }
  size_t size() const { return size_; }
  bool geographic() const { return geographic_; }

 private:
  mdc_obs* h_ = nullptr;
  size_t size_ = 0;
  bool geographic_ = false;
};

}  // namespace metada::backends::cuda
