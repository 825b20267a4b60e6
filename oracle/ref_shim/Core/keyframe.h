// oracle/ref_shim/Core/keyframe.h — TEST INFRASTRUCTURE ONLY: KeyFrame is a Frame (Core/keyframe.h:11); nothing else of it is used
// by Odometry/ransac.cpp.
#pragma once
#include "frame.h"
class KeyFrame : public Frame {};
