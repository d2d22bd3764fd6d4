// pcl_min.h — layout-identical stand-ins for the few PCL / Eigen types that cross the hot-path
// boundary, used ONLY when the real headers are absent (this image has no PCL).
// With PCL installed, include <pcl/point_types.h> / <pcl/point_cloud.h> instead: the adapter code
// only relies on  sizeof(pcl::PointXYZRGB) == 32, sizeof(pcl::Normal) == 32, x/y/z at offset 0,
// normal_x.. at 0 and curvature at 16, and `cloud->points` being contiguous
// (include/Path_Generate.h:32-33,66-68 of the reference).
#pragma once
#if defined(__has_include)
#if __has_include(<pcl/point_types.h>) && __has_include(<pcl/point_cloud.h>) && !defined(PPP_FORCE_PCL_MIN)
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#define PPP_HAVE_REAL_PCL 1
#endif
#endif

#ifndef PPP_HAVE_REAL_PCL
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <vector>

namespace Eigen {
struct Vector3f {
  float v[3];
  Vector3f() : v{0, 0, 0} {}
  Vector3f(float a, float b, float c) : v{a, b, c} {}
  float& operator[](int i) { return v[i]; }
  float operator[](int i) const { return v[i]; }
};
struct Vector3d {
  double v[3];
  Vector3d() : v{0, 0, 0} {}
  Vector3d(double a, double b, double c) : v{a, b, c} {}
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};
}  // namespace Eigen

namespace pcl {
struct alignas(16) PointXYZRGB {
  float x, y, z, pad_ = 1.0f;
  union {
    struct { uint8_t b, g, r, a; };
    float rgb;
    uint32_t rgba;
  };
  float pad2_[3];
  PointXYZRGB() : x(0), y(0), z(0), rgba(0xff000000u), pad2_{0, 0, 0} {}
};
struct alignas(16) Normal {
  float normal_x, normal_y, normal_z, pad_ = 0.0f;
  float curvature;
  float pad2_[3];
  Normal() : normal_x(0), normal_y(0), normal_z(0), curvature(0), pad2_{0, 0, 0} {}
};
static_assert(sizeof(PointXYZRGB) == 32, "pcl::PointXYZRGB layout");
static_assert(sizeof(Normal) == 32, "pcl::Normal layout");

template <typename T>
struct AlignedAlloc {
  typedef T value_type;
  AlignedAlloc() {}
  template <typename U> AlignedAlloc(const AlignedAlloc<U>&) {}
  T* allocate(size_t n) { return static_cast<T*>(aligned_alloc(16, ((n * sizeof(T) + 15) / 16) * 16)); }
  void deallocate(T* p, size_t) { free(p); }
  template <typename U> bool operator==(const AlignedAlloc<U>&) const { return true; }
  template <typename U> bool operator!=(const AlignedAlloc<U>&) const { return false; }
};

template <typename PointT>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  std::vector<PointT, AlignedAlloc<PointT>> points;
  uint32_t width = 0, height = 1;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
};
}  // namespace pcl
#endif  // !PPP_HAVE_REAL_PCL
