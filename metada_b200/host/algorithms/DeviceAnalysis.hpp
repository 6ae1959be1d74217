#pragma once
// Device analysis hook: what an algorithm policy calls instead of its CPU body when the backend
// tag provides a device path.  The policies in this directory (LETKF.hpp, ETKF.hpp, EnKF.hpp) keep
// the reference's class interface (LETKF.hpp:47-48,63,125; ETKF.hpp:86-87,100,185;
// EnKF.hpp:105-106,139,261,287) and delegate Analyse() here.
#include <vector>

#include "CudaApi.hpp"
#include "Ensemble.hpp"
#include "Observation.hpp"

namespace metada::framework::device {

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
    ptrs.push_back(ensemble.GetMember(m).template getDataPtr<double>());
  }
  dev->upload(ptrs);
  // a geometry with 2-D coordinate arrays (WRFGeometry::unstaggered_info(), the arrays
  // IdentityObsOperator.hpp:488-526 searches) gives the device store its geography
  if constexpr (requires { g.unstaggered_info().latitude_2d; g.unstaggered_info().longitude_2d; g.unstaggered_info().vertical_coords; }) {
    const auto& info = g.unstaggered_info();
    if (info.has_2d_coords())
      dev->setGeography(std::vector<double>(info.latitude_2d.begin(), info.latitude_2d.end()),
                        std::vector<double>(info.longitude_2d.begin(), info.longitude_2d.end()),
                        std::vector<double>(info.vertical_coords.begin(), info.vertical_coords.end()));
  }
  return dev;
}

/** Writes the analysed members back in place (LETKF.hpp:240-242 / ETKF.hpp:172-176 semantics). */
template <typename BackendTag>
void downloadEnsemble(backends::cuda::DeviceEnsemble& dev, Ensemble<BackendTag>& ensemble) {
  std::vector<double*> ptrs;
  for (size_t m = 0; m < ensemble.Size(); ++m) ptrs.push_back(ensemble.GetMember(m).template getDataPtr<double>());
  dev.download(ptrs);
}

}  // namespace metada::framework::device
