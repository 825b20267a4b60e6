// oracle/ref_shim/pcl/filters/statistical_outlier_removal.h — TEST INFRASTRUCTURE ONLY: see voxel_grid.h of this directory.
#pragma once
#include <cstdlib>
#include <pcl/point_cloud.h>
namespace pcl {
template <typename PointT> class StatisticalOutlierRemoval {
public:
    template <typename P> void setInputCloud(const P&) {}
    void setMeanK(int) {}
    void setStddevMulThresh(double) {}
    void filter(PointCloud<PointT>&) { std::abort(); }
};
}  // namespace pcl
