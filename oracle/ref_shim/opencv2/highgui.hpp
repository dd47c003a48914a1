// TEST INFRASTRUCTURE ONLY: part of the OpenCV stand-in (see opencv2/core.hpp); everything lives in core.hpp.
#pragma once
#include <opencv2/core.hpp>
