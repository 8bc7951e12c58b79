// Map ingestion and configuration — the step BEFORE the hot path in the reference:
//   TRGPlanner::setParams        src/planner/trg_planner.cpp:103-129   (yaml-cpp)
//   TRGPlanner::loadPrebuiltMap  src/planner/trg_planner.cpp:76-101    (pcl::io::loadPCDFile + pcl::VoxelGrid)
// yaml-cpp and PCL are absent from this image; the YAML subset of config/*.yaml and the PCD file
// format (ascii / binary / binary_compressed, fields x y z as 4-byte floats) are read by the small
// parsers in map_io.cpp. The voxel filter runs on the device (include/trgb_kernels.h, K8).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace trg_b200 {

// TRGPlanner::paramStruct as filled by setParams, defaults included (planner.cpp:107-128)
struct PlannerParams {
  bool  isVerbose     = true;
  float graph_rate    = 1.0f;
  float planning_rate = 1.0f;
  bool  isPreMap      = false;
  std::string preMapPath;
  bool  isVoxelize    = false;
  float VoxelSize     = 0.1f;
  bool  isPreGraph    = false;
  std::string preGraphPath;
  bool  isUpdate                 = false;
  float expandDist               = 0.6f;
  float robotSize                = 0.3f;
  int   sampleNum                = 20;
  float heightThreshold          = 0.15f;
  float collisionThreshold       = 0.2f;
  float updateCollisionThreshold = 0.2f;
  float safetyFactor             = 1.0f;
  float goal_tolerance           = 0.8f;
};

// throws std::runtime_error when the file cannot be read
PlannerParams load_params_yaml(const std::string& config_path);

// packed xyz triples; throws std::runtime_error on a missing file / unsupported layout
std::vector<float> load_pcd_xyz(const std::string& path);
void save_pcd_xyz(const std::string& path, const float* xyz, int64_t n, bool binary);

}  // namespace trg_b200
