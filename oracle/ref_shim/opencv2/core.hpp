// oracle/ref_shim/opencv2/core.hpp — TEST INFRASTRUCTURE ONLY: everything the reference needs lives in opencv2/core/core.hpp of this shim.
#pragma once
#include "core/core.hpp"
