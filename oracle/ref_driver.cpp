// ref_driver.cpp -- runs the reference's OWN, UNMODIFIED headers (LETKF.hpp, ETKF.hpp, EnKF.hpp,
// Ensemble/State/Observation/ObsOperator adapters, Simple backend, IdentityObsOperator, Location)
// from /root/reference on a JSON config and dumps full-precision results for tests/golden.
//
// TEST INFRASTRUCTURE ONLY (oracle/): builds into oracle/_ref (git-ignored), against the Eigen-API
// shim in oracle/eigen_shim because Eigen is absent from this image (see oracle/README.md).
//
//   ref_driver hx             cfg.json out.bin   H(x_j) for every member, obs values/variances,
//                                                local-obs counts via Location::distance_to
//   ref_driver letkf          cfg.json out.bin   LETKF<SimpleBackendTag>::Analyse() exactly as shipped
//   ref_driver letkf_snapshot cfg.json out.bin   same LETKF.hpp, but the ObsOperator BACKEND is a
//                                                caching wrapper so H is evaluated on the background
//                                                ensemble only (snapshot semantics)
//   ref_driver etkf | enkf    cfg.json out.bin   ETKF / EnKF <SimpleBackendTag>::Analyse()
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "EnKF.hpp"
#include "LWEnKF.hpp"
#include "ETKF.hpp"
#include "LETKF.hpp"

#include "ApplicationContext.hpp"
#include "Config.hpp"
#include "ControlVariableBackend.hpp"
#include "Ensemble.hpp"
#include "Geometry.hpp"
#include "IdentityControlVariableBackend.hpp"
#include "ObsOperator.hpp"
#include "Observation.hpp"
#include "SimpleBackendTraits.hpp"

namespace fwk = metada::framework;
using SimpleTag = metada::traits::SimpleBackendTag;

// ---- snapshot variant: identical traits, except H results are cached per state object ----------
namespace refsnap {
template <typename StateBackend, typename ObsBackend, typename CvBackend>
class SnapshotObsOperator {
  using Inner = metada::backends::common::obsoperator::IdentityObsOperator<StateBackend, ObsBackend, CvBackend>;

 public:
  SnapshotObsOperator() = delete;
  SnapshotObsOperator(const SnapshotObsOperator&) = delete;
  SnapshotObsOperator& operator=(const SnapshotObsOperator&) = delete;
  SnapshotObsOperator(SnapshotObsOperator&&) noexcept = default;
  SnapshotObsOperator& operator=(SnapshotObsOperator&&) noexcept = default;
  template <typename ConfigBackend>
  explicit SnapshotObsOperator(const ConfigBackend& config) : inner_(config) {}
  template <typename ConfigBackend>
  void initialize(const ConfigBackend& config) { inner_.initialize(config); }
  bool isInitialized() const { return inner_.isInitialized(); }
  std::vector<double> apply(const StateBackend& state, const ObsBackend& obs) const {
    auto it = cache_.find(&state);
    if (it == cache_.end()) it = cache_.emplace(&state, inner_.apply(state, obs)).first;
    return it->second;
  }
  const std::vector<std::string>& getRequiredStateVars() const { return inner_.getRequiredStateVars(); }
  const std::vector<std::string>& getRequiredObsVars() const { return inner_.getRequiredObsVars(); }

 private:
  Inner inner_;
  mutable std::map<const StateBackend*, std::vector<double>> cache_;
};
struct SnapshotTag {};
}  // namespace refsnap

namespace metada::traits {
template <>
struct BackendTraits<refsnap::SnapshotTag> : BackendTraits<SimpleBackendTag> {
  using Base = BackendTraits<SimpleBackendTag>;
  using ObsOperatorBackend =
      refsnap::SnapshotObsOperator<Base::StateBackend, Base::ObservationBackend, Base::ControlVariableBackend>;
};
}  // namespace metada::traits

static void write_i64(std::ofstream& f, int64_t v) { f.write(reinterpret_cast<const char*>(&v), 8); }
static void write_vec(std::ofstream& f, const std::vector<double>& v) {
  write_i64(f, static_cast<int64_t>(v.size()));
  f.write(reinterpret_cast<const char*>(v.data()), static_cast<std::streamsize>(v.size() * 8));
}

template <typename Tag>
static std::vector<double> dump_members(fwk::Ensemble<Tag>& ens) {
  std::vector<double> out;
  for (size_t m = 0; m < ens.Size(); ++m) {
    auto& mem = ens.GetMember(m);
    const double* p = mem.template getDataPtr<double>();
    out.insert(out.end(), p, p + mem.size());
  }
  return out;
}

template <typename Tag>
static int run(const std::string& mode, int argc, char** argv, const std::string& out_path) {
  auto context = fwk::ApplicationContext<Tag>(argc, argv);
  auto& config = context.getConfig();
  fwk::Geometry<Tag> geometry(config.GetSubsection("geometry"));
  fwk::Ensemble<Tag> ensemble(config.GetSubsection("ensemble"), geometry);
  fwk::Observation<Tag> observations(config.GetSubsection("observations"));
  auto control_backend = std::make_shared<fwk::IdentityControlVariableBackend<Tag>>();
  fwk::ObsOperator<Tag> obs_operator(config.GetSubsection("obs_operator"), *control_backend);

  std::ofstream f(out_path, std::ios::binary);
  const int64_t k = static_cast<int64_t>(ensemble.Size());
  const int64_t n = static_cast<int64_t>(ensemble.GetMember(0).size());
  const int64_t P = static_cast<int64_t>(observations.size());
  write_i64(f, k); write_i64(f, n); write_i64(f, P);

  if (mode == "hx") {
    for (int64_t m = 0; m < k; ++m) write_vec(f, obs_operator.apply(ensemble.GetMember(m), observations));
    write_vec(f, observations.getObservationValues());
    write_vec(f, observations.getCovariance());
    // observation grid coordinates as the reference parsed them
    std::vector<double> ox, oy, oz;
    std::vector<fwk::Location> locs;
    for (const auto& op : observations) {
      auto [i, j, kk] = op.location.getGridCoords();
      ox.push_back(i); oy.push_back(j); oz.push_back(kk);
      locs.push_back(op.location);
    }
    write_vec(f, ox); write_vec(f, oy); write_vec(f, oz);
    // local-observation counts per grid point, LETKF.hpp:159-165 with the config's radius
    const double radius = config.GetSubsection("analysis").Get("localization_radius").asFloat();
    std::vector<double> counts;
    const auto* geom = ensemble.GetMember(0).geometry();
    for (const auto& gp : *geom) {
      int c = 0;
      for (const auto& l : locs)
        if (gp.distance_to(l) <= radius) ++c;
      counts.push_back(c);
    }
    write_vec(f, counts);
    std::vector<double> rad{radius, static_cast<double>(config.GetSubsection("analysis").Get("inflation").asFloat())};
    write_vec(f, rad);
    // Ensemble::RecomputeMean (Ensemble.hpp:105-114)
    ensemble.RecomputeMean();
    const double* mp = ensemble.Mean().template getDataPtr<double>();
    write_vec(f, std::vector<double>(mp, mp + n));
    return 0;
  }
  if (mode == "letkf" || mode == "letkf_snapshot") {
    if (mode == "letkf_snapshot")   // pre-warm the H cache on the BACKGROUND ensemble
      for (int64_t m = 0; m < k; ++m) (void)obs_operator.apply(ensemble.GetMember(m), observations);
    fwk::LETKF<Tag> letkf(ensemble, observations, obs_operator, config.GetSubsection("analysis"));
    letkf.Analyse();
    write_vec(f, dump_members(ensemble));
    const double* mp = ensemble.Mean().template getDataPtr<double>();
    write_vec(f, std::vector<double>(mp, mp + n));
    return 0;
  }
  if (mode == "etkf") {
    fwk::ETKF<Tag> etkf(ensemble, observations, obs_operator, config.GetSubsection("analysis"));
    etkf.Analyse();
    write_vec(f, dump_members(ensemble));
    return 0;
  }
  if (mode == "enkf") {
    fwk::EnKF<Tag> enkf(ensemble, observations, obs_operator, config);
    enkf.Analyse();
    write_vec(f, dump_members(ensemble));
    // the N(0,1) draws EnKF.hpp:346-356 consumed (same generator, same fixed seed, same order)
    std::random_device rd;   // -> metada_fixed_random_device via ref_fixed_seed.hpp
    std::mt19937 gen(rd());
    std::normal_distribution<double> dist(0.0, 1.0);
    std::vector<double> Z(static_cast<size_t>(P * k));
    for (int64_t i = 0; i < P; ++i)
      for (int64_t j = 0; j < k; ++j) Z[static_cast<size_t>(i * k + j)] = dist(gen);
    write_vec(f, Z);
    auto r = enkf.getAnalysisResults();
    write_vec(f, {r.innovation_norm, r.background_spread, r.analysis_spread, r.max_kalman_gain,
                  r.min_kalman_gain, r.condition_number, r.inflation_factor});
    return 0;
  }
  if (mode == "lwenkf") {
    fwk::LWEnKF<Tag> lw(ensemble, observations, obs_operator, config);
    lw.Analyse();
    write_vec(f, dump_members(ensemble));
    // the N(0,1) draws LWEnKF.hpp:666-679 consumed (same generator, same fixed seed, same order)
    std::random_device rd;   // -> metada_fixed_random_device via ref_fixed_seed.hpp
    std::mt19937 gen(rd());
    std::normal_distribution<double> dist(0.0, 1.0);
    std::vector<double> Z(static_cast<size_t>(P * k));
    for (int64_t i = 0; i < P; ++i)
      for (int64_t j = 0; j < k; ++j) Z[static_cast<size_t>(i * k + j)] = dist(gen);
    write_vec(f, Z);
    auto r = lw.getAnalysisResults();
    write_vec(f, {r.innovation_norm, r.background_spread, r.analysis_spread, r.max_kalman_gain, r.min_kalman_gain,
                  r.condition_number, r.max_weight, r.min_weight, r.weight_variance});
    return 0;
  }
  std::cerr << "unknown mode " << mode << std::endl;
  return 2;
}

int main(int argc, char** argv) {
  if (argc != 4) {
    std::cerr << "usage: ref_driver <hx|letkf|letkf_snapshot|etkf|enkf|lwenkf> <config.json> <out.bin>" << std::endl;
    return 2;
  }
  const std::string mode = argv[1], out = argv[3];
  char* av[2] = {argv[0], argv[2]};
  try {
    if (mode == "letkf_snapshot") return run<refsnap::SnapshotTag>(mode, 2, av, out);
    return run<SimpleTag>(mode, 2, av, out);
  } catch (const std::exception& e) {
    std::cerr << "ref_driver failed: " << e.what() << std::endl;
    return 1;
  }
}
