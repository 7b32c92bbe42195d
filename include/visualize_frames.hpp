// visualize_frames.hpp -- alternate entry of the visualiser for a device-resident simulation.
//
// visualize::UpdateVisualization (reference include/visualize.hpp, src/visualize.cpp:150-219) receives the 15
// FP64 fields and, before its first OpenCV call, narrows 12 quantities to CV_32F matrices and reads 19 values at
// 9 sample points.  With the fields living on a GPU that narrowing is done there (plbm_frames_begin, plbm.h) and
// only its results travel to the host.  A maintainer adopts it by splitting the three Visualization* functions
// of visualize.cpp after their conversion loops:
//
//     cv::Mat mat_n_e(NY, NX, CV_32F, const_cast<float*>(mats[VF_RHO_E]));   // instead of the at<float>() loop
//     ... normalize_and_color / wrap_with_label / hconcat / write as before ...
//     ts_ux_e[t][p] = series[VS_UX_E][p];                                     // instead of indexing the fields
//
// The unchanged UpdateVisualization path stays available (LBmethod::Run_simulation).
#pragma once

namespace visualize {

enum FrameId { VF_RHO_E = 0, VF_RHO_I, VF_RHO_Q, VF_UX_E, VF_UY_E, VF_UE_MAG, VF_UX_I, VF_UY_I, VF_UI_MAG, VF_T_E, VF_T_I, VF_T_N, VF_COUNT };
enum SeriesId { VS_UX_E = 0, VS_UY_E, VS_UE_MAG, VS_UX_I, VS_UY_I, VS_UI_MAG, VS_UX_N, VS_UY_N, VS_UN_MAG,
                VS_T_E, VS_T_I, VS_T_N, VS_RHO_E, VS_RHO_I, VS_RHO_N, VS_RHO_Q, VS_EX, VS_EY, VS_E_MAG, VS_COUNT };

// mats[k]: NY rows of NX floats, mat(y, x) exactly as the reference's cv::Mat(NY, NX, CV_32F) after its conversion
// loop; series[q][p]: quantity q at sample point p (order of InitVisualization, src/visualize.cpp:72-80).
void UpdateVisualizationFrames(int t, int NX, int NY, const float* const mats[VF_COUNT], const double series[VS_COUNT][9]);

} // namespace visualize
