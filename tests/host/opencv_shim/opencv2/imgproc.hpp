/* tests/host/opencv_shim: see opencv.hpp */
#pragma once
#include "opencv.hpp"
