// Minimal stand-ins for pcl::PointXYZ / pcl::PointCloud (common.h:62-63 of the reference:
// `using PtsDefault = pcl::PointXYZ; using PointCloudPtr = std::shared_ptr<pcl::PointCloud<PtsDefault>>`),
// used only when PCL is absent. Layout matches PCL: a 16-byte, 16-aligned point (x, y, z, pad).
#pragma once
#include <cstddef>
#include <memory>
#include <vector>

namespace pcl {

struct alignas(16) PointXYZ {
  float x = 0.f, y = 0.f, z = 0.f, data_c = 1.f;
  PointXYZ() = default;
  PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_), data_c(1.f) {}
};

template <typename PointT>
class PointCloud {
 public:
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  std::vector<PointT> points;
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); }
  void push_back(const PointT& p) { points.push_back(p); }
  void resize(std::size_t n) { points.resize(n); }
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
};

}  // namespace pcl
