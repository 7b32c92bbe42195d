// streaming.cpp -- streaming::{Stream, Streaming*, ThermalStreaming*} with the reference's
// signatures (include/streaming.hpp) over plbm_host_stream.
#include "streaming.hpp"
#include "plbm.h"

#include <stdexcept>
#include <string>

namespace streaming {

namespace {
// scatter the three arrays into temp_* on the device, then swap like the reference (src/streaming.cpp:56-58)
void stream3(std::vector<double>& a, std::vector<double>& b, std::vector<double>& c,
             std::vector<double>& ta, std::vector<double>& tb, std::vector<double>& tc,
             int NX, int NY, int bc, const char* what)
{
    const double* in[3] = { a.data(), b.data(), c.data() };
    double* out[3] = { ta.data(), tb.data(), tc.data() };
    if (plbm_host_stream(NX, NY, bc, in, out)) throw std::runtime_error(std::string(what) + ": " + plbm_last_error());
    a.swap(ta); b.swap(tb); c.swap(tc);
}
}

void StreamingPeriodic(std::vector<double>& f_e, std::vector<double>& f_i, std::vector<double>& f_n,
                       std::vector<double>& temp_e, std::vector<double>& temp_i, std::vector<double>& temp_n,
                       const std::array<int, Q>&, const std::array<int, Q>&, const int NX, const int NY)
{
    stream3(f_e, f_i, f_n, temp_e, temp_i, temp_n, NX, NY, PLBM_BC_PERIODIC, "streaming::StreamingPeriodic");
}

void ThermalStreamingPeriodic(std::vector<double>& g_e, std::vector<double>& g_i, std::vector<double>& g_n,
                              std::vector<double>& temp_e, std::vector<double>& temp_i, std::vector<double>& temp_n,
                              const std::array<int, Q>&, const std::array<int, Q>&, const int NX, const int NY)
{
    stream3(g_e, g_i, g_n, temp_e, temp_i, temp_n, NX, NY, PLBM_BC_PERIODIC, "streaming::ThermalStreamingPeriodic");
}

void StreamingBounceBack(std::vector<double>& f_e, std::vector<double>& f_i, std::vector<double>& f_n,
                         std::vector<double>& temp_e, std::vector<double>& temp_i, std::vector<double>& temp_n,
                         const std::array<int, Q>&, const std::array<int, Q>&, const int NX, const int NY)
{
    stream3(f_e, f_i, f_n, temp_e, temp_i, temp_n, NX, NY, PLBM_BC_BOUNCEBACK, "streaming::StreamingBounceBack");
}

void ThermalStreamingBounceBack(std::vector<double>& g_e, std::vector<double>& g_i, std::vector<double>& g_n,
                                std::vector<double>& temp_e, std::vector<double>& temp_i, std::vector<double>& temp_n,
                                const std::array<int, Q>&, const std::array<int, Q>&, const int NX, const int NY)
{
    stream3(g_e, g_i, g_n, temp_e, temp_i, temp_n, NX, NY, PLBM_BC_BOUNCEBACK, "streaming::ThermalStreamingBounceBack");
}

void Stream(std::vector<double>& f_e, std::vector<double>& f_i, std::vector<double>& f_n,
            std::vector<double>& temp_e, std::vector<double>& temp_i, std::vector<double>& temp_n,
            std::vector<double>& g_e, std::vector<double>& g_i, std::vector<double>& g_n,
            const std::array<int, Q>& cx, const std::array<int, Q>& cy,
            const int NX, const int NY, const BCType type)
{
    switch (type) {                                              // reference src/streaming.cpp:18-29
    case BCType::Periodic:
        StreamingPeriodic(f_e, f_i, f_n, temp_e, temp_i, temp_n, cx, cy, NX, NY);
        ThermalStreamingPeriodic(g_e, g_i, g_n, temp_e, temp_i, temp_n, cx, cy, NX, NY);
        break;
    case BCType::BounceBack:
        StreamingBounceBack(f_e, f_i, f_n, temp_e, temp_i, temp_n, cx, cy, NX, NY);
        ThermalStreamingBounceBack(g_e, g_i, g_n, temp_e, temp_i, temp_n, cx, cy, NX, NY);
        break;
    default:
        throw std::runtime_error("Tipo di streaming non supportato.");
    }
}

} // namespace streaming
