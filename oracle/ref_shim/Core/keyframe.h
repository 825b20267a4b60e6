// oracle/ref_shim/Core/keyframe.h — TEST INFRASTRUCTURE ONLY: KeyFrame is a Frame (Core/keyframe.h:11) with its own copy of the image
// bounds; IsInImage restated from Core/keyframe.cpp:424-427.  Nothing else of it is used by Odometry/ransac.cpp or Features/matcher.cpp.
#pragma once
#include "frame.h"
class KeyFrame : public Frame {
public:
    bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }
    float mnMinX = 0.f, mnMaxX = 640.f, mnMinY = 0.f, mnMaxY = 480.f;
};
