// tests/host/vis_dump.cpp -- test double for the reference's OpenCV renderer: the three
// visualize:: entry points LBmethod::Run_simulation calls, writing the fields they receive as
// raw float64 (order of the parameter list) for the steps listed in $PLBM_DUMP_STEPS.
#include "visualize.hpp"
#include "visualize_frames.hpp"

#include <cstdio>
#include <cstdlib>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>

namespace {
std::set<int> steps() {
    std::set<int> s;
    if (const char* e = std::getenv("PLBM_DUMP_STEPS")) { std::stringstream ss(e); std::string tok; while (std::getline(ss, tok, ',')) if (!tok.empty()) s.insert(std::atoi(tok.c_str())); }
    return s;
}
}

namespace visualize {
void InitVisualization(const int, const int, const int) {}
void CloseVisualization() {}
void UpdateVisualization(const int t, const int NX, const int NY,
    const std::vector<double>& ux_e,  const std::vector<double>& uy_e, const std::vector<double>& ux_i,  const std::vector<double>& uy_i,
    const std::vector<double>& ux_n,  const std::vector<double>& uy_n, const std::vector<double>& T_e,   const std::vector<double>& T_i,
    const std::vector<double>& T_n, const std::vector<double>& rho_e, const std::vector<double>& rho_i, const std::vector<double>& rho_n,
    const std::vector<double>& rho_q, const std::vector<double>& Ex,    const std::vector<double>& Ey)
{
    static const std::set<int> want = steps();
    const char* dir = std::getenv("PLBM_DUMP_DIR");
    if (!dir || !want.count(t)) return;
    char name[512];
    std::snprintf(name, sizeof(name), "%s/fields_t%05d.f64", dir, t);
    FILE* fp = std::fopen(name, "wb");
    if (!fp) throw std::runtime_error(std::string("cannot open ") + name);
    const std::vector<double>* all[15] = { &ux_e, &uy_e, &ux_i, &uy_i, &ux_n, &uy_n, &T_e, &T_i, &T_n, &rho_e, &rho_i, &rho_n, &rho_q, &Ex, &Ey };
    for (auto* f : all) std::fwrite(f->data(), sizeof(double), static_cast<size_t>(NX) * NY, fp);
    std::fclose(fp);
}

// alternate entry: the 12 CV_32F matrices as raw float32, then the 19 x 9 series as float64
void UpdateVisualizationFrames(int t, int NX, int NY, const float* const mats[VF_COUNT], const double series[VS_COUNT][9])
{
    static const std::set<int> want = steps();
    const char* dir = std::getenv("PLBM_DUMP_DIR");
    if (!dir || !want.count(t)) return;
    char name[512];
    std::snprintf(name, sizeof(name), "%s/frames_t%05d.bin", dir, t);
    FILE* fp = std::fopen(name, "wb");
    if (!fp) throw std::runtime_error(std::string("cannot open ") + name);
    for (int k = 0; k < VF_COUNT; ++k) std::fwrite(mats[k], sizeof(float), static_cast<size_t>(NX) * NY, fp);
    std::fwrite(series, sizeof(double), static_cast<size_t>(VS_COUNT) * 9, fp);
    std::fclose(fp);
}
}
