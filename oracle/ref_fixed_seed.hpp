// Force-included (-include) when building oracle/_ref: EnKF.hpp:346-347 seeds its mt19937 from
// std::random_device, so its perturbed observations are unrepeatable.  Without touching the
// reference source, every later use of the NAME random_device becomes this fixed-seed stand-in;
// the driver regenerates the identical N(0,1) stream with the same libstdc++ generator.
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include <random>
namespace std {
struct metada_fixed_random_device {
  unsigned int operator()() { return 20261017u; }
};
}  // namespace std
#define random_device metada_fixed_random_device
