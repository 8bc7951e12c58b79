// ORACLE — TEST INFRASTRUCTURE ONLY. Stand-in for the two PCL containers trg.cpp touches
// (pcl::PointXYZ, pcl::PointCloud<T>: common.h:62-63, trg.cpp:183,201,301,318,742). PCL is absent
// from this image; no PCL algorithm runs on the path (VoxelGrid / loadPCDFile are in the caller).
#ifndef ORACLE_SHIM_PCL_POINT_TYPES_H_
#define ORACLE_SHIM_PCL_POINT_TYPES_H_
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>
// the real PCL / OpenCV / nlohmann headers drag these in; trg.cpp relies on it
#include <algorithm>
#include <fstream>
#include <functional>
#include <iomanip>
#include <limits>
#include <queue>
#include <sstream>

namespace pcl {
struct alignas(16) PointXYZ {  // 16 bytes like the real one (x, y, z + padding)
  float x, y, z, pad_;
  PointXYZ() : x(0.f), y(0.f), z(0.f), pad_(1.f) {}
  PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_), pad_(1.f) {}
};

template <typename PointT>
class PointCloud {
 public:
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 0;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() {
    points.clear();
    width = height = 0;
  }
  void push_back(const PointT& p) {
    points.push_back(p);
    width = static_cast<std::uint32_t>(points.size());
    height = 1;
  }
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
};
}  // namespace pcl
#endif
