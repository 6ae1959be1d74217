#pragma once
// Device analysis hook: what an algorithm policy calls instead of its CPU body when the backend
// tag provides a device path.  The policies in this directory (LETKF.hpp, ETKF.hpp, EnKF.hpp, LWEnKF.hpp) keep
// the reference's class interface (LETKF.hpp:47-48,63,125; ETKF.hpp:86-87,100,185;
// EnKF.hpp:105-106,139,261,287) and delegate Analyse() here.
//
// The ensemble is DEVICE-RESIDENT across analyses (the reference's Ensemble owns k host States,
// Ensemble.hpp:189, and has no ensemble-level backend hook; the hook therefore sits in the state
// backend): acquireResident() finds or creates the [col][lev][member] store the members are linked
// to and uploads only the members whose host arrays may have been written since the last upload;
// publishResident() declares the device copy the truth after an analysis WITHOUT downloading --
// a member comes back when (and only if) somebody reads it on the host (CudaState::syncHost), and
// the ensemble mean comes from the device (mdc_ens_mean, n doubles instead of k n).
#include <memory>
#include <utility>
#include <vector>

#include "CudaApi.hpp"
#include "Ensemble.hpp"
#include "Observation.hpp"

namespace metada::framework::device {

template <typename Geo>
void attachGeography(backends::cuda::DeviceEnsemble& dev, const Geo& g) {
  // a geometry with 2-D coordinate arrays (WRFGeometry::unstaggered_info(), the arrays
  // IdentityObsOperator.hpp:488-526 searches) gives the device store its geography
  if constexpr (requires { g.unstaggered_info().latitude_2d; g.unstaggered_info().longitude_2d; g.unstaggered_info().vertical_coords; }) {
    const auto& info = g.unstaggered_info();
    if (info.has_2d_coords())
      dev.setGeography(std::vector<double>(info.latitude_2d.begin(), info.latitude_2d.end()),
                       std::vector<double>(info.longitude_2d.begin(), info.longitude_2d.end()),
                       std::vector<double>(info.vertical_coords.begin(), info.vertical_coords.end()));
  }
}

/** Gathers the members' host arrays (State::getDataPtr<double>, [lev][y][x]) into one device store. */
template <typename BackendTag>
std::unique_ptr<backends::cuda::DeviceEnsemble> uploadEnsemble(Ensemble<BackendTag>& ensemble) {
  const auto* geometry = ensemble.GetMember(0).geometry();
  if (!geometry) throw std::runtime_error("Geometry pointer is null in device analysis");
  const auto& g = geometry->backend();
  const int k = static_cast<int>(ensemble.Size());
  auto dev = std::make_unique<backends::cuda::DeviceEnsemble>(g.x_dim(), g.y_dim(), g.z_dim(), k);
  std::vector<const double*> ptrs;
  for (int m = 0; m < k; ++m) {
    if (ensemble.GetMember(m).size() != dev->pointsPerMember())
      throw std::runtime_error("ensemble member size does not match the geometry");
    ptrs.push_back(std::as_const(ensemble).GetMember(m).template getDataPtr<double>());
  }
  dev->upload(ptrs);
  attachGeography(*dev, g);
  return dev;
}

/** Writes the analysed members back in place (LETKF.hpp:240-242 / ETKF.hpp:172-176 semantics). */
template <typename BackendTag>
void downloadEnsemble(backends::cuda::DeviceEnsemble& dev, Ensemble<BackendTag>& ensemble) {
  std::vector<double*> ptrs;
  for (size_t m = 0; m < ensemble.Size(); ++m) ptrs.push_back(ensemble.GetMember(m).template getDataPtr<double>());
  dev.download(ptrs);
}

/** The device store this ensemble's members are linked to (created and linked on first use), with
 *  every member whose host array may be newer than the device copy uploaded (runs of consecutive
 *  members go in one call). */
template <typename BackendTag>
std::shared_ptr<backends::cuda::DeviceEnsemble> acquireResident(Ensemble<BackendTag>& ensemble) {
  using backends::cuda::DeviceEnsemble;
  using backends::cuda::ResidentLink;
  const auto* geometry = ensemble.GetMember(0).geometry();
  if (!geometry) throw std::runtime_error("Geometry pointer is null in device analysis");
  const auto& g = geometry->backend();
  const int k = static_cast<int>(ensemble.Size());
  std::shared_ptr<DeviceEnsemble> store;
  if (const auto& l0 = ensemble.GetMember(0).backend().resident()) store = l0->store;
  bool linked = store && store->members() == k && store->nx() == static_cast<int>(g.x_dim()) &&
                store->ny() == static_cast<int>(g.y_dim()) && store->nz() == static_cast<int>(g.z_dim());
  for (int m = 0; linked && m < k; ++m) {
    const auto& l = ensemble.GetMember(m).backend().resident();
    linked = l && l->store == store && l->member == m;
  }
  if (!linked) {
    // (members still linked to another store hold the newest values there: bring them home first)
    for (int m = 0; m < k; ++m) ensemble.GetMember(m).backend().syncHost();
    store = std::make_shared<DeviceEnsemble>(g.x_dim(), g.y_dim(), g.z_dim(), k);
    attachGeography(*store, g);
    for (int m = 0; m < k; ++m) {
      if (ensemble.GetMember(m).size() != store->pointsPerMember())
        throw std::runtime_error("ensemble member size does not match the geometry");
      auto link = std::make_shared<ResidentLink>();
      link->store = store;
      link->member = m;
      ensemble.GetMember(m).backend().attachResident(std::move(link));
    }
  }
  for (int m = 0; m < k;) {
    if (!ensemble.GetMember(m).backend().resident()->device_stale) { ++m; continue; }
    const int m0 = m;
    std::vector<const double*> ptrs;
    for (; m < k && ensemble.GetMember(m).backend().resident()->device_stale; ++m) {
      auto& b = ensemble.GetMember(m).backend();
      b.syncHost();                                           // (cannot be stale on both sides; cheap no-op)
      ptrs.push_back(b.hostShadow());
    }
    store->uploadMembers(m0, ptrs);
    for (int i = m0; i < m; ++i) ensemble.GetMember(i).backend().resident()->device_stale = false;
  }
  return store;
}

/** After an analysis that updated the store in place: the device holds the truth, nothing is copied. */
template <typename BackendTag>
void publishResident(Ensemble<BackendTag>& ensemble) {
  for (size_t m = 0; m < ensemble.Size(); ++m) {
    const auto& l = ensemble.GetMember(m).backend().resident();
    l->host_stale = true;
    l->device_stale = false;
  }
}

/** upload -> body(store) -> results stay on the device (resident) or come straight back (not resident). */
template <typename BackendTag, typename Body>
void analyseOnDevice(Ensemble<BackendTag>& ensemble, bool resident, Body&& body) {
  if (resident) {
    auto dev = acquireResident(ensemble);
    bool have_mean = true;
    try { (void)std::as_const(ensemble).Mean(); } catch (const std::runtime_error&) { have_mean = false; }
    // Ensemble creates its mean state only inside RecomputeMean(): run it once while the host values are
    // current (no transfers); afterwards the mean is Ensemble::RecomputeMean on the device (Ensemble.hpp:105-114:
    // same sum order, same multiplication by 1.0 / k -- bit-identical, tests/test_gpu_golden.py), n doubles
    // over PCIe instead of k n
    if (!have_mean) ensemble.RecomputeMean();
    body(*dev);
    publishResident(ensemble);
    dev->mean(ensemble.Mean().template getDataPtr<double>());
  } else {
    auto dev = uploadEnsemble(ensemble);
    body(*dev);
    downloadEnsemble(*dev, ensemble);
    ensemble.RecomputeMean();                                  // LETKF.hpp:116
  }
}

}  // namespace metada::framework::device
