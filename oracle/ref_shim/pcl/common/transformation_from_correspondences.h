// oracle/ref_shim/pcl/common/transformation_from_correspondences.h — TEST INFRASTRUCTURE ONLY.
// pcl::TransformationFromCorrespondences (PCL 1.8, common/include/pcl/common/transformation_from_correspondences.h and its
// impl/*.hpp) is not in this image.  The stand-in keeps the (point, point, weight) triples in the order add() receives them and hands
// them to the oracle's restatement of the published algorithm (incremental weighted mean / covariance recurrence, 3x3 SVD, reflection
// fix-up; oracle/match_ransac_oracle.cpp: Tfc behind orc_tfc_transform) — one statement of the PCL arithmetic for both sides.
#pragma once
#include <vector>
#include <Eigen/Core>
namespace pcl {
class TransformationFromCorrespondences {
public:
    void reset() { p_.clear(); q_.clear(); w_.clear(); }
    void add(const Eigen::Vector3f& point, const Eigen::Vector3f& corresponding_point, float weight = 1.0f)
    {
        for (int i = 0; i < 3; ++i) { p_.push_back(point(i)); q_.push_back(corresponding_point(i)); }
        w_.push_back(weight);
    }
    Eigen::Affine3f getTransformation()
    {
        float T[16];
        orc_tfc_transform(p_.data(), q_.data(), w_.data(), (int)w_.size(), T);
        Eigen::Matrix4f m;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m(i, j) = T[4 * i + j];
        return Eigen::Affine3f(m);
    }
private:
    std::vector<float> p_, q_, w_;
};
}  // namespace pcl
