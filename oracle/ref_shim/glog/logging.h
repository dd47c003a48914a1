// TEST INFRASTRUCTURE ONLY: header/config.h:28 includes glog for Ceres; main.cpp:16 calls InitGoogleLogging.
#pragma once
namespace google { inline void InitGoogleLogging(const char*) {} }
