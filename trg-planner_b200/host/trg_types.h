// Value types at the TRG API boundary: the real Eigen / PCL when installed, the lite
// stand-ins otherwise (same names, so callers written against the reference compile as-is).
#pragma once
#include <memory>

#if defined(__has_include)
#if __has_include(<Eigen/Core>)
#include <Eigen/Core>
#define TRG_B200_HAVE_EIGEN 1
#endif
#if __has_include(<pcl/point_types.h>) && __has_include(<pcl/point_cloud.h>)
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#define TRG_B200_HAVE_PCL 1
#endif
#endif
#ifndef TRG_B200_HAVE_EIGEN
#include "compat/eigen_lite.h"
#endif
#ifndef TRG_B200_HAVE_PCL
#include "compat/pcl_lite.h"
#endif

// common.h:62-63
using PtsDefault    = pcl::PointXYZ;
using PointCloudPtr = std::shared_ptr<pcl::PointCloud<PtsDefault>>;
