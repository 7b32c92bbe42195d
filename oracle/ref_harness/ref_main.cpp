// oracle/ref_harness/ref_main.cpp -- TEST INFRASTRUCTURE ONLY.
// Driver for the UNMODIFIED reference solver (sources compiled in place from /root/reference by
// oracle/Makefile).  It plays the role of /root/reference/src/main_plasma.cpp:7-97 with the
// hard-coded parameters turned into command-line flags, and either calls the reference's own
// LBmethod::Run_simulation() or replays its loop body (plasma.cpp:476-523) phase by phase with
// timers.  `private` is opened only in this TU to reach UpdateMacro/ComputeEquilibrium and the
// population vectors.
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <iostream>
#include <string>
#include <utility>
#include <vector>
#include <omp.h>

#define private public
#include "plasma.hpp"
#undef private
#include "vis_hook.hpp"

namespace poisson { const std::vector<double>& oracle_phi(); }

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void write_vec(FILE* fp, const std::vector<double>& v) { std::fwrite(v.data(), sizeof(double), v.size(), fp); }

int main(int argc, char** argv)
{
    int NX = 200, NY = 200, NSTEPS = 200, threads = 1, poisson_i = 3, bc_i = 0;
    int Z_ion = 1, A_ion = 1;
    double n_e = 1e11, n_n = 1e18, T_e = 1e4, T_i = 300, T_n = 300, Ex_SI = 1e-2, Ey_SI = 0.0, omega = 1.8;
    bool phases = false, pops = false;
    int periodic_solver = 0, calls = 1;          // --periodic-solver 1|2|4: call poisson::SolvePoisson_{GS,SOR,9point}_Periodic directly
    std::string out, dump, rhoq_path;
    for (int i = 1; i < argc; ++i) {
        auto is = [&](const char* k) { return std::strcmp(argv[i], k) == 0; };
        auto next = [&]() { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", argv[i]); std::exit(2); } return argv[++i]; };
        if (is("--nx")) NX = std::atoi(next());
        else if (is("--ny")) NY = std::atoi(next());
        else if (is("--steps")) NSTEPS = std::atoi(next());
        else if (is("--threads")) threads = std::atoi(next());
        else if (is("--poisson")) poisson_i = std::atoi(next());
        else if (is("--bc")) bc_i = std::atoi(next());
        else if (is("--omega")) omega = std::atof(next());
        else if (is("--ex")) Ex_SI = std::atof(next());
        else if (is("--ey")) Ey_SI = std::atof(next());
        else if (is("--te")) T_e = std::atof(next());
        else if (is("--ti")) T_i = std::atof(next());
        else if (is("--tn")) T_n = std::atof(next());
        else if (is("--ne")) n_e = std::atof(next());
        else if (is("--nn")) n_n = std::atof(next());
        else if (is("--zion")) Z_ion = std::atoi(next());
        else if (is("--aion")) A_ion = std::atoi(next());
        else if (is("--out")) out = next();
        else if (is("--dump")) dump = next();
        else if (is("--pops")) pops = true;
        else if (is("--phases")) phases = true;
        else if (is("--periodic-solver")) periodic_solver = std::atoi(next());
        else if (is("--calls")) calls = std::atoi(next());
        else if (is("--rhoq")) rhoq_path = next();
        else { std::fprintf(stderr, "unknown flag %s\n", argv[i]); return 2; }
    }
    ref_hook::Config& hook = ref_hook::config();
    hook.out_dir = out;
    if (!out.empty()) std::filesystem::create_directories(out);
    if (dump == "all") hook.dump_all = true;
    else {
        size_t pos = 0;
        while (pos < dump.size()) {
            size_t c = dump.find(',', pos);
            if (c == std::string::npos) c = dump.size();
            if (c > pos) hook.dump_steps.insert(std::atoi(dump.substr(pos, c - pos).c_str()));
            pos = c + 1;
        }
    }
    const auto ptype = static_cast<poisson::PoissonType>(poisson_i);
    const auto btype = static_cast<streaming::BCType>(bc_i);

    if (periodic_solver) {
        // The *_Periodic solvers (src/poisson.cpp:146-211, 283-354, 487-546) are public but never called by the reference's own
        // loop.  They work on the file-static phi, which only poisson::SolvePoisson sizes (call_once, :34-41): one call with
        // PoissonType::NONE does that (and zeroes E), then the solver under test runs `calls` times on rho_q read from --rhoq
        // (warm start, like successive time steps).  phi after the last call goes to <out>/phi_periodic.f64.
        omp_set_num_threads(threads);
        std::vector<double> rho_q((size_t)NX * NY), Ex((size_t)NX * NY), Ey((size_t)NX * NY);
        FILE* fr = std::fopen(rhoq_path.c_str(), "rb");
        if (!fr || std::fread(rho_q.data(), sizeof(double), rho_q.size(), fr) != rho_q.size()) { std::perror("--rhoq"); return 1; }
        std::fclose(fr);
        poisson::SolvePoisson(Ex, Ey, rho_q, NX, NY, omega, poisson::PoissonType::NONE, streaming::BCType::Periodic);
        for (int k = 0; k < calls; ++k) {
            if (periodic_solver == 1) poisson::SolvePoisson_GS_Periodic(rho_q, NX, NY);
            else if (periodic_solver == 2) poisson::SolvePoisson_SOR_Periodic(rho_q, NX, NY, omega);
            else poisson::SolvePoisson_9point_Periodic(rho_q, NX, NY);
        }
        const std::string name = out + "/phi_periodic.f64";
        FILE* fp = std::fopen(name.c_str(), "wb");
        if (!fp) { std::perror(name.c_str()); return 1; }
        write_vec(fp, poisson::oracle_phi());
        std::fclose(fp);
        std::printf("{\"periodic_solver\": %d, \"calls\": %d, \"nx\": %d, \"ny\": %d}\n", periodic_solver, calls, NX, NY);
        return 0;
    }

    const double t0 = now_s();
    LBmethod lb(NSTEPS, NX, NY, static_cast<size_t>(threads), Z_ion, A_ion, Ex_SI, Ey_SI, T_e, T_i, T_n, n_e, n_n,
                ptype, btype, omega);
    const double t1 = now_s();
    double ph[5] = {0, 0, 0, 0, 0};
    if (!phases) {
        lb.Run_simulation();
    } else {
        omp_set_num_threads(threads);
        visualize::InitVisualization(NX, NY, NSTEPS);
        for (int t = 0; t < NSTEPS; ++t) {
            double a = now_s();
            lb.UpdateMacro();
            double b = now_s(); ph[0] += b - a; a = b;
            lb.ComputeEquilibrium();
            b = now_s(); ph[1] += b - a; a = b;
            collisions::Collide(lb.g_e, lb.g_i, lb.g_n, lb.g_eq_e, lb.g_eq_i, lb.g_eq_n,
                                lb.g_eq_e_i, lb.g_eq_e_n, lb.g_eq_i_n, lb.g_eq_i_e, lb.g_eq_n_e, lb.g_eq_n_i,
                                lb.f_e, lb.f_i, lb.f_n, lb.f_eq_e, lb.f_eq_i, lb.f_eq_n,
                                lb.f_eq_e_i, lb.f_eq_e_n, lb.f_eq_i_n, lb.f_eq_i_e, lb.f_eq_n_e, lb.f_eq_n_i,
                                lb.rho_e, lb.rho_i, lb.rho_n, lb.ux_e, lb.uy_e, lb.ux_i, lb.uy_i, lb.ux_n, lb.uy_n,
                                lb.Ex, lb.Ey, lb.q_e, lb.q_i, lb.m_e, lb.m_i, lb.temp_e, lb.temp_i, lb.temp_n,
                                LBmethod::cx, LBmethod::cy, LBmethod::w, NX, NY, lb.Kb, lb.cs2);
            b = now_s(); ph[2] += b - a; a = b;
            streaming::Stream(lb.f_e, lb.f_i, lb.f_n, lb.temp_e, lb.temp_i, lb.temp_n, lb.g_e, lb.g_i, lb.g_n,
                              LBmethod::cx, LBmethod::cy, NX, NY, btype);
            b = now_s(); ph[3] += b - a; a = b;
            poisson::SolvePoisson(lb.Ex, lb.Ey, lb.rho_q, NX, NY, omega, ptype, btype);
            b = now_s(); ph[4] += b - a;
            visualize::UpdateVisualization(t, NX, NY, lb.ux_e, lb.uy_e, lb.ux_i, lb.uy_i, lb.ux_n, lb.uy_n,
                                           lb.T_e, lb.T_i, lb.T_n, lb.rho_e, lb.rho_i, lb.rho_n, lb.rho_q, lb.Ex, lb.Ey);
        }
        visualize::CloseVisualization();
    }
    const double t2 = now_s();

    if (pops && !out.empty()) {
        const std::string name = out + "/pops_final.f64";
        FILE* fp = std::fopen(name.c_str(), "wb");
        if (!fp) { std::perror(name.c_str()); return 1; }
        write_vec(fp, lb.f_e); write_vec(fp, lb.f_i); write_vec(fp, lb.f_n);
        write_vec(fp, lb.g_e); write_vec(fp, lb.g_i); write_vec(fp, lb.g_n);
        std::fclose(fp);
    }
    const double loop_s = t2 - t1;
    std::printf("{\"nx\": %d, \"ny\": %d, \"steps\": %d, \"threads\": %d, \"poisson\": %d, \"bc\": %d, "
                "\"ctor_s\": %.6f, \"loop_s\": %.6f, \"mlups\": %.6f, "
                "\"phase_s\": {\"update_macro\": %.6f, \"equilibrium\": %.6f, \"collide\": %.6f, \"stream\": %.6f, \"poisson\": %.6f}, "
                "\"units\": {\"cs2\": %.17g, \"Kb\": %.17g, \"m_e\": %.17g, \"m_i\": %.17g, \"m_n\": %.17g, \"q_e\": %.17g, \"q_i\": %.17g, "
                "\"rho_e_init\": %.17g, \"rho_i_init\": %.17g, \"rho_n_init\": %.17g, \"T_e_init\": %.17g, \"T_i_init\": %.17g, \"T_n_init\": %.17g, "
                "\"Ex_ext\": %.17g, \"Ey_ext\": %.17g}}\n",
                NX, NY, NSTEPS, threads, poisson_i, bc_i, t1 - t0, loop_s,
                loop_s > 0 ? (double)NX * NY * NSTEPS / loop_s / 1e6 : 0.0,
                ph[0], ph[1], ph[2], ph[3], ph[4],
                lb.cs2, lb.Kb, lb.m_e, lb.m_i, lb.m_n, lb.q_e, lb.q_i, lb.rho_e_init, lb.rho_i_init, lb.rho_n_init,
                lb.T_e_init, lb.T_i_init, lb.T_n_init, lb.Ex_ext, lb.Ey_ext);
    return 0;
}
