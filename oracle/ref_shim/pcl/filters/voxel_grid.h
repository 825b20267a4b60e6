// oracle/ref_shim/pcl/filters/voxel_grid.h — TEST INFRASTRUCTURE ONLY: pcl::VoxelGrid exists so that Core/frame.cpp compiles; dense-cloud
// filtering is not on the hot path and stops the process if a test ever reaches it.
#pragma once
#include <cstdlib>
#include <pcl/point_cloud.h>
namespace pcl {
template <typename PointT> class VoxelGrid {
public:
    void setLeafSize(float, float, float) {}
    template <typename P> void setInputCloud(const P&) {}
    void filter(PointCloud<PointT>&) { std::abort(); }
};
}  // namespace pcl
