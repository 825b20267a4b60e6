// oracle/ref_shim/pcl/point_types.h — TEST INFRASTRUCTURE ONLY: pcl::PointXYZ as far as Odometry/ransac.cpp uses it.
#pragma once
namespace pcl {
struct PointXYZ {
    float x, y, z, pad;                       // PCL pads the point to 16 bytes (data[3] = 1)
    PointXYZ() : x(0), y(0), z(0), pad(1.f) {}
    PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_), pad(1.f) {}
};
struct PointNormal : PointXYZ { float normal_x = 0, normal_y = 0, normal_z = 0, curvature = 0; };
struct PointXYZRGB : PointXYZ { unsigned char b, g, r, a; PointXYZRGB() : b(0), g(0), r(0), a(255) {} };
}  // namespace pcl
