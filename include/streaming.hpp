// streaming.hpp -- public surface of the reference's streaming module (reference
// include/streaming.hpp:10-60), implemented on the B200 through include/plbm.h.
// Every function borrows caller-owned host vectors in the reference layout and returns its result
// the way the reference does: the streamed populations end up in f_*/g_* and temp_* holds the
// previous contents (the reference scatters into temp_* and swaps, src/streaming.cpp:50-58).
#pragma once

#include "utils.hpp"

#include <array>
#include <vector>

namespace streaming {

enum class BCType {
    Periodic,
    BounceBack
};

// dispatcher (reference src/streaming.cpp:13-30); throws std::runtime_error for an unknown BCType
void Stream(std::vector<double>& f_e, std::vector<double>& f_i, std::vector<double>& f_n,
            std::vector<double>& temp_e, std::vector<double>& temp_i, std::vector<double>& temp_n,
            std::vector<double>& g_e, std::vector<double>& g_i, std::vector<double>& g_n,
            const std::array<int, Q>& cx, const std::array<int, Q>& cy,
            const int NX, const int NY, const BCType type);

void StreamingPeriodic(std::vector<double>& f_e, std::vector<double>& f_i, std::vector<double>& f_n,
                       std::vector<double>& temp_e, std::vector<double>& temp_i, std::vector<double>& temp_n,
                       const std::array<int, Q>& cx, const std::array<int, Q>& cy,
                       const int NX, const int NY);

void StreamingBounceBack(std::vector<double>& f_e, std::vector<double>& f_i, std::vector<double>& f_n,
                         std::vector<double>& temp_e, std::vector<double>& temp_i, std::vector<double>& temp_n,
                         const std::array<int, Q>& cx, const std::array<int, Q>& cy,
                         const int NX, const int NY);

void ThermalStreamingPeriodic(std::vector<double>& g_e, std::vector<double>& g_i, std::vector<double>& g_n,
                              std::vector<double>& temp_e, std::vector<double>& temp_i, std::vector<double>& temp_n,
                              const std::array<int, Q>& cx, const std::array<int, Q>& cy,
                              const int NX, const int NY);

void ThermalStreamingBounceBack(std::vector<double>& g_e, std::vector<double>& g_i, std::vector<double>& g_n,
                                std::vector<double>& temp_e, std::vector<double>& temp_i, std::vector<double>& temp_n,
                                const std::array<int, Q>& cx, const std::array<int, Q>& cy,
                                const int NX, const int NY);

} // namespace streaming
