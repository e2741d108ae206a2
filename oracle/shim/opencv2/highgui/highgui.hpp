// empty on purpose: see core/core.hpp (no-op OpenCV stand-in for the headless oracle build)
#pragma once
#include "opencv2/core/core.hpp"
