// oracle/ref_shim/pcl/point_cloud.h — TEST INFRASTRUCTURE ONLY: pcl::PointCloud as far as Odometry/ransac.cpp uses it
// (a vector of points behind a shared pointer; PCL 1.8's Ptr is a boost::shared_ptr, std::shared_ptr here).
#pragma once
#include <memory>
#include <vector>
#include <boost/make_shared.hpp>      // the real PCL headers bring boost's smart pointers with them
namespace pcl {
template <typename PointT> class PointCloud {
public:
    typedef PointT PointType;
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
    std::vector<PointT> points;
    size_t size() const { return points.size(); }
};
}  // namespace pcl
