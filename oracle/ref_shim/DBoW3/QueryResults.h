// oracle/ref_shim/DBoW3/QueryResults.h — TEST INFRASTRUCTURE ONLY: nothing of DBoW3's query results is used on the hot path.
#pragma once
#include "DBoW3.h"
