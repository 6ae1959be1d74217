#pragma once
// State backend of the CUDA backend.  Satisfies framework::StateBackendImpl
// (StateConcepts.hpp:47-84) plus the un-concepted members the adapters use: at(Location),
// geometry(), operator[] (State.hpp:385-386; IdentityObsOperator.hpp:246,265).
//
// A CudaState is the HOST view of one member, layout [lev][y][x] (what State::getDataPtr<double>()
// hands to ETKF.hpp:172-176 / EnKF.hpp:230-234).  For an analysis the members are gathered into one
// device-resident [col][lev][member] store (CudaApi.hpp: DeviceEnsemble), which STAYS on the device
// across analyses: the host array is a shadow, synchronised lazily through a ResidentLink -- every
// accessor below downloads this member first if the device holds newer values (host_stale), and every
// accessor that can write marks the device copy stale so that the next analysis uploads this member
// again (DeviceAnalysis.hpp: acquireResident / publishResident).
// File formats: whitespace-separated text in [lev][y][x] order (SimpleState.hpp:248-263 for
// z_dim == 1); saveToFile keeps the reference's fixed 6-decimal format (SimpleState.hpp:162-188).
#include <cmath>
#include <filesystem>
#include <fstream>
#include <iomanip>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "CudaApi.hpp"
#include "CudaGeometry.hpp"
#include "Location.hpp"

namespace metada::backends::cuda {

class CudaState {
 public:
  CudaState() = delete;
  CudaState(const CudaState&) = delete;
  CudaState& operator=(const CudaState&) = delete;
  CudaState(CudaState&&) noexcept = default;
  CudaState& operator=(CudaState&&) = delete;
  ~CudaState() = default;

  template <typename ConfigBackend>
  CudaState(const ConfigBackend& config, const CudaGeometry& geometry) : geometry_(geometry) {
    try {
      auto v = config.Get("variables");
      variable_names_ = v.isString() ? std::vector<std::string>{v.asString()} : v.asVectorString();
    } catch (...) {
      variable_names_ = {"state"};
    }
    readFromFile(config.Get("file").asString());
  }

  void* getData() { touch(); return data_.data(); }
  const void* getData() const { syncHost(); return data_.data(); }
  const std::vector<std::string>& getVariableNames() const { return variable_names_; }
  size_t size() const { return data_.size(); }
  const CudaGeometry& geometry() const { return geometry_; }

  void zero() {
    if (link_) { link_->host_stale = false; link_->device_stale = true; }   // (nothing worth downloading)
    std::fill(data_.begin(), data_.end(), 0.0);
  }
  void add(const CudaState& o) { same(o, "add"); touch(); o.syncHost(); for (size_t i = 0; i < data_.size(); ++i) data_[i] += o.data_[i]; }
  void subtract(const CudaState& o) { same(o, "subtract"); touch(); o.syncHost(); for (size_t i = 0; i < data_.size(); ++i) data_[i] -= o.data_[i]; }
  void multiply(double s) { touch(); for (auto& v : data_) v *= s; }
  double dot(const CudaState& o) const {
    same(o, "compute dot product of");
    syncHost(); o.syncHost();
    double r = 0.0;
    for (size_t i = 0; i < data_.size(); ++i) r += data_[i] * o.data_[i];
    return r;
  }
  double norm() const { return std::sqrt(dot(*this)); }
  bool equals(const CudaState& o) const { syncHost(); o.syncHost(); return data_ == o.data_; }
  template <typename IncrementBackend>
  void addIncrement(const IncrementBackend& inc) {
    touch();
    const auto v = inc.getData();
    for (size_t i = 0; i < data_.size() && i < v.size(); ++i) data_[i] += v[i];
  }

  std::unique_ptr<CudaState> clone() const { syncHost(); return std::unique_ptr<CudaState>(new CudaState(*this, 0)); }

  // ---- device residency (see the header comment); used by framework::device::acquireResident / publishResident
  const std::shared_ptr<ResidentLink>& resident() const { return link_; }
  void attachResident(std::shared_ptr<ResidentLink> link) { link_ = std::move(link); }
  /** Host array without synchronisation: for the transfers themselves. */
  double* hostShadow() { return data_.data(); }
  /** Brings the host shadow up to date (one member's download) if the device holds newer values. */
  void syncHost() const {
    if (link_ && link_->host_stale) {
      link_->store->downloadMember(link_->member, const_cast<double*>(data_.data()));
      link_->host_stale = false;
    }
  }

  void saveToFile(const std::string& filename) const {
    syncHost();
    std::filesystem::path p(filename);
    if (!p.parent_path().empty() && !std::filesystem::exists(p.parent_path()))
      std::filesystem::create_directories(p.parent_path());
    std::ofstream f(filename);
    if (!f.is_open()) throw std::runtime_error("Could not open file for writing: " + filename);
    const size_t nx = geometry_.x_dim(), rows = static_cast<size_t>(geometry_.y_dim()) * geometry_.z_dim();
    for (size_t r = 0; r < rows; ++r) {
      for (size_t x = 0; x < nx; ++x) {
        f << std::fixed << std::setprecision(6) << std::setw(12) << data_[r * nx + x];
        if (x + 1 < nx) f << " ";
      }
      f << "\n";
    }
  }

  double& at(const framework::Location& loc) { touch(); return data_[index(loc)]; }
  const double& at(const framework::Location& loc) const { syncHost(); return data_[index(loc)]; }
  double& operator[](size_t i) { if (i >= data_.size()) throw std::out_of_range("Index out of range"); touch(); return data_[i]; }
  const double& operator[](size_t i) const { if (i >= data_.size()) throw std::out_of_range("Index out of range"); syncHost(); return data_[i]; }

 private:
  // (a clone is a plain host state: no link)
  CudaState(const CudaState& o, int) : data_(o.data_), variable_names_(o.variable_names_), geometry_(o.geometry_) {}
  /** Host access that may write: up-to-date shadow first, then the device copy counts as stale. */
  void touch() {
    syncHost();
    if (link_) link_->device_stale = true;
  }
  void same(const CudaState& o, const char* what) const {
    if (data_.size() != o.data_.size()) throw std::runtime_error(std::string("Cannot ") + what + " states of different sizes");
  }
  size_t index(const framework::Location& loc) const {
    auto [i, j, k] = loc.getGridCoords();
    return (static_cast<size_t>(k) * geometry_.y_dim() + j) * geometry_.x_dim() + i;
  }
  void readFromFile(const std::string& filename) {
    std::ifstream f(filename);
    if (!f) throw std::runtime_error("Cannot open state file: " + filename);
    std::string line;
    while (std::getline(f, line)) {
      std::istringstream iss(line);
      double v;
      while (iss >> v) data_.push_back(v);
    }
    if (data_.size() != geometry_.size())
      throw std::runtime_error("State file " + filename + " holds " + std::to_string(data_.size()) +
                               " values, geometry has " + std::to_string(geometry_.size()));
  }

  std::vector<double> data_;
  std::vector<std::string> variable_names_;
  const CudaGeometry& geometry_;
  std::shared_ptr<ResidentLink> link_;
};

/** Host vector space over the grid: IncrementBackend and ControlVariableBackend of the CUDA
 *  backend (IncrementConcepts.hpp:45-73, ControlVariableConcepts.hpp:52-86).  The ensemble filters
 *  never touch it; it exists because ObsOperator.hpp:10-12,85-100 needs the types. */
class CudaIncrement {
 public:
  explicit CudaIncrement(const CudaGeometry& g) : data_(g.size(), 0.0), geometry_(&g) {}
  void zero() { std::fill(data_.begin(), data_.end(), 0.0); }
  void scale(double a) { for (auto& v : data_) v *= a; }
  void axpy(double a, const CudaIncrement& o) { for (size_t i = 0; i < data_.size(); ++i) data_[i] += a * o.data_[i]; }
  double dot(const CudaIncrement& o) const { double r = 0; for (size_t i = 0; i < data_.size(); ++i) r += data_[i] * o.data_[i]; return r; }
  double norm() const { return std::sqrt(dot(*this)); }
  CudaIncrement& operator+=(const CudaIncrement& o) { axpy(1.0, o); return *this; }
  CudaIncrement& operator-=(const CudaIncrement& o) { axpy(-1.0, o); return *this; }
  CudaIncrement& operator*=(double s) { scale(s); return *this; }
  CudaIncrement& operator/=(double s) { scale(1.0 / s); return *this; }
  size_t size() const { return data_.size(); }
  const CudaGeometry& geometry() const { return *geometry_; }
  std::vector<double> getData() const { return data_; }
  void setFromVector(const std::vector<double>& v) { data_.assign(v.begin(), v.end()); data_.resize(geometry_->size(), 0.0); }
  void randomize() {
    unsigned long long s = 0x9E3779B97F4A7C15ull;
    for (auto& v : data_) { s = s * 6364136223846793005ull + 1442695040888963407ull; v = static_cast<double>(s >> 11) / 9007199254740992.0 - 0.5; }
  }
  template <typename StateBackend>
  void transferFromState(const StateBackend& st) {
    const double* p = static_cast<const double*>(st.getData());
    std::copy(p, p + data_.size(), data_.begin());
  }

 private:
  std::vector<double> data_;
  const CudaGeometry* geometry_;
};

}  // namespace metada::backends::cuda
