#pragma once
// Geometry backend of the CUDA backend: a regular x_dim x y_dim x z_dim grid.
// Satisfies framework::GeometryBackendImpl (GeometryConcepts.hpp:43-96).  Iteration yields GRID
// Locations level-major (k outer, then y, then x): for z_dim == 1 this is SimpleGeometry's order
// (SimpleGeometry.hpp:50-54), for z_dim > 1 the WRF order (WRFGeometryIterator.hpp:113-116).
#include <cstddef>
#include <iterator>
#include <limits>
#include <stdexcept>

#include "Location.hpp"

namespace metada::backends::cuda {

class CudaGeometry;

class CudaGeometryIterator {
 public:
  using iterator_category = std::forward_iterator_tag;
  using value_type = framework::Location;
  using difference_type = std::ptrdiff_t;
  using pointer = const value_type*;
  using reference = value_type;

  CudaGeometryIterator() = default;
  CudaGeometryIterator(int nx, int ny, size_t idx) : nx_(nx), ny_(ny), idx_(idx) {}
  value_type operator*() const { return at(idx_); }
  struct Arrow {
    value_type v;
    const value_type* operator->() const { return &v; }
  };
  Arrow operator->() const { return Arrow{at(idx_)}; }
  CudaGeometryIterator& operator++() { ++idx_; return *this; }
  CudaGeometryIterator operator++(int) { auto t = *this; ++idx_; return t; }
  bool operator==(const CudaGeometryIterator& o) const { return idx_ == o.idx_; }
  bool operator!=(const CudaGeometryIterator& o) const { return idx_ != o.idx_; }
  friend difference_type operator-(const CudaGeometryIterator& a, const CudaGeometryIterator& b) {
    return static_cast<difference_type>(a.idx_) - static_cast<difference_type>(b.idx_);
  }

 private:
  value_type at(size_t idx) const {
    const size_t plane = static_cast<size_t>(nx_) * ny_;
    const int k = static_cast<int>(idx / plane);
    const size_t r = idx % plane;
    return framework::Location(static_cast<int>(r % nx_), static_cast<int>(r / nx_), k);
  }
  int nx_ = 1, ny_ = 1;
  size_t idx_ = 0;
};

class CudaGeometry {
 public:
  using value_type = framework::Location;
  using reference = value_type;
  using const_reference = const value_type;
  using pointer = value_type*;
  using const_pointer = const value_type*;
  using size_type = std::size_t;
  using difference_type = std::ptrdiff_t;
  using iterator = CudaGeometryIterator;
  using const_iterator = CudaGeometryIterator;

  CudaGeometry() = delete;
  CudaGeometry(const CudaGeometry&) = delete;
  CudaGeometry& operator=(const CudaGeometry&) = delete;
  CudaGeometry(CudaGeometry&&) noexcept = default;
  CudaGeometry& operator=(CudaGeometry&&) noexcept = default;

  /** config: x_dim, y_dim (as SimpleGeometry.hpp:42-48) and optional z_dim (default 1). */
  template <typename ConfigBackend>
  explicit CudaGeometry(const ConfigBackend& config) {
    x_dim_ = config.Get("x_dim").asInt();
    y_dim_ = config.Get("y_dim").asInt();
    try {
      z_dim_ = config.Get("z_dim").asInt();
    } catch (...) {
      z_dim_ = 1;
    }
    if (x_dim_ <= 0 || y_dim_ <= 0 || z_dim_ <= 0)
      throw std::invalid_argument("CudaGeometry: x_dim, y_dim and z_dim must be positive");
  }
  CudaGeometry(int x_dim, int y_dim, int z_dim) : x_dim_(x_dim), y_dim_(y_dim), z_dim_(z_dim) {
    if (x_dim_ <= 0 || y_dim_ <= 0 || z_dim_ <= 0)
      throw std::invalid_argument("CudaGeometry: x_dim, y_dim and z_dim must be positive");
  }

  iterator begin() { return iterator(x_dim_, y_dim_, 0); }
  iterator end() { return iterator(x_dim_, y_dim_, size()); }
  const_iterator begin() const { return const_iterator(x_dim_, y_dim_, 0); }
  const_iterator end() const { return const_iterator(x_dim_, y_dim_, size()); }
  const_iterator cbegin() const { return begin(); }
  const_iterator cend() const { return end(); }

  size_type size() const { return static_cast<size_type>(x_dim_) * y_dim_ * z_dim_; }
  bool empty() const { return size() == 0; }
  size_type max_size() const { return std::numeric_limits<size_type>::max(); }

  reference operator[](size_type idx) { return *iterator(x_dim_, y_dim_, idx); }
  const_reference operator[](size_type idx) const { return *iterator(x_dim_, y_dim_, idx); }
  reference at(size_type idx) { check(idx); return (*this)[idx]; }
  const_reference at(size_type idx) const { check(idx); return (*this)[idx]; }
  reference front() { return (*this)[0]; }
  const_reference front() const { return (*this)[0]; }
  reference back() { return (*this)[size() - 1]; }
  const_reference back() const { return (*this)[size() - 1]; }

  CudaGeometry clone() const { return CudaGeometry(x_dim_, y_dim_, z_dim_); }

  int x_dim() const { return x_dim_; }
  int y_dim() const { return y_dim_; }
  int z_dim() const { return z_dim_; }

 private:
  void check(size_type idx) const {
    if (idx >= size()) throw std::out_of_range("Grid point index out of range");
  }
  int x_dim_ = 0, y_dim_ = 0, z_dim_ = 1;
};

}  // namespace metada::backends::cuda
