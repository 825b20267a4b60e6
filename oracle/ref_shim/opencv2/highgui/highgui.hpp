// oracle/ref_shim/opencv2/highgui/highgui.hpp — TEST INFRASTRUCTURE ONLY: everything the reference needs lives in opencv2/core/core.hpp of this shim.
#pragma once
#include "../../opencv2/core/core.hpp"
