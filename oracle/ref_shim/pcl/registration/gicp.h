// oracle/ref_shim/pcl/registration/gicp.h — TEST INFRASTRUCTURE ONLY: the member type Odometry/generalizedicp.h declares; GICP itself is
// a third-party iterative solver outside the hot path and is never run here.
#pragma once
#include <Eigen/Core>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
namespace pcl {
template <typename PointSource, typename PointTarget> class GeneralizedIterativeClosestPoint {};
}  // namespace pcl
