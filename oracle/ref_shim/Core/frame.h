// oracle/ref_shim/Core/frame.h — TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's Core/frame.h (which pulls in DBoW3, PCL filters and the whole map graph): the one member of Frame
// that Odometry/ransac.cpp reads — the camera-frame 3D point of every keypoint (Core/frame.h:87) — plus the descriptor matrix.
#pragma once
#include <vector>
#include <opencv2/opencv.hpp>
class Frame {
public:
    virtual ~Frame() {}
    std::vector<cv::Point3f> mvKeys3Dc;
    cv::Mat mDescriptors;
};
