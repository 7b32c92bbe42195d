// visualize.hpp -- the consumer at the boundary.  LBmethod::Run_simulation hands 15 host fields
// to visualize::UpdateVisualization every step (reference src/plasma.cpp:516-522); this repo does
// not re-implement the OpenCV renderer, it only has to keep the reference's src/visualize.cpp
// compiling unchanged against this header (declarations of reference include/visualize.hpp:12-98).
// Without OpenCV headers only the three entry points the solver calls are declared, so the solver
// and its tests build anywhere; link any implementation of them.
#pragma once

#include "utils.hpp"

#include <array>
#include <filesystem>
#include <string>
#include <utility>
#include <vector>

#if defined(__has_include)
#  if __has_include(<opencv2/opencv.hpp>)
#    include <opencv2/opencv.hpp>
#    define PLBM_HAVE_OPENCV 1
#  endif
#endif

namespace visualize {

void InitVisualization(const int NX, const int NY, const int T);

void UpdateVisualization(const int t, const int NX, const int NY,
    const std::vector<double>& ux_e,  const std::vector<double>& uy_e,
    const std::vector<double>& ux_i,  const std::vector<double>& uy_i,
    const std::vector<double>& ux_n,  const std::vector<double>& uy_n,
    const std::vector<double>& T_e,   const std::vector<double>& T_i,
    const std::vector<double>& T_n,
    const std::vector<double>& rho_e, const std::vector<double>& rho_i,
    const std::vector<double>& rho_n, const std::vector<double>& rho_q,
    const std::vector<double>& Ex,    const std::vector<double>& Ey);

void CloseVisualization();

#ifdef PLBM_HAVE_OPENCV
// ---- what the reference's renderer (src/visualize.cpp) expects to find declared -----------------
static constexpr int P = 9;                                   // sample points: centre + 8 around
extern std::array<std::pair<int, int>, P> sample_points;

using Series = std::vector<std::array<double, P>>;            // [time][point]
extern Series ts_ux_e, ts_uy_e, ts_ue_mag;
extern Series ts_ux_i, ts_uy_i, ts_ui_mag;
extern Series ts_ux_n, ts_uy_n, ts_un_mag;
extern Series ts_T_e, ts_T_i, ts_T_n;
extern Series ts_rho_e, ts_rho_i, ts_rho_n, ts_rho_q;
extern Series ts_Ex, ts_Ey, ts_E_mag;

extern cv::VideoWriter video_writer_density;
extern cv::VideoWriter video_writer_velocity;
extern cv::VideoWriter video_writer_temperature;

cv::Mat PlotTimeSeriesWithOpenCV(const Series& data, const std::string& title);

void VisualizationDensity(const int NX, const int NY,
    const std::vector<double>& rho_e, const std::vector<double>& rho_i, const std::vector<double>& rho_q);
void VisualizationVelocity(const int NX, const int NY,
    const std::vector<double>& ux_e, const std::vector<double>& uy_e,
    const std::vector<double>& ux_i, const std::vector<double>& uy_i);
void VisualizationTemperature(const int NX, const int NY,
    const std::vector<double>& T_e, const std::vector<double>& T_i, const std::vector<double>& T_n);

cv::Mat normalize_and_color(const cv::Mat& src, const double vmin, const double vmax);
cv::Mat wrap_with_label(const cv::Mat& img, const std::string& label);
#endif

} // namespace visualize
