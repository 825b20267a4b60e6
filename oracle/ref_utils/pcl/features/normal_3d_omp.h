// stand-in: Utils/utils.cpp includes this header for AddNormal, which the TUM-file comparison never calls (see ../utils_stubs.h)
#pragma once
