/*
 * oracle/shim/opencv2/opencv.hpp -- TEST INFRASTRUCTURE ONLY.
 * Parse-only stand-in so that /root/reference/include/visualize.hpp:10 (pulled in by
 * plasma.hpp:7) compiles in an image without OpenCV C++ headers.  The reference's
 * src/visualize.cpp is NOT compiled against this; oracle/ref_harness/vis_hook.cpp provides the
 * three visualize:: entry points instead.
 */
#pragma once
#include <array>
#include <string>
#include <utility>
#include <vector>
namespace cv {
class Mat {};
class VideoWriter {};
}
