// plbm_host_api.cu -- stateless per-phase entry points of include/plbm.h on HOST arrays.
// Each call uploads its inputs, runs one kernel of phases.cu, and downloads the outputs.
#include "../../include/plbm.h"

#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "phases.h"
#include "exact_math.cuh"

using namespace plbm;

int plbm_set_error(const char* msg);   // plbm_api.cu

namespace {

struct Pool {   // device scratch released on scope exit
    std::vector<void*> bufs;
    cudaError_t err = cudaSuccess;
    ~Pool() { for (void* p : bufs) cudaFree(p); }
    double* up(const double* host, size_t n)
    {
        double* d = alloc(n);
        if (d && host && err == cudaSuccess) err = cudaMemcpy(d, host, sizeof(double) * n, cudaMemcpyHostToDevice);
        return d;
    }
    double* alloc(size_t n)
    {
        void* d = nullptr;
        if (err == cudaSuccess) err = cudaMalloc(&d, sizeof(double) * n);
        if (d) bufs.push_back(d);
        return (double*)d;
    }
    bool down(double* host, const double* dev, size_t n)
    {
        if (err == cudaSuccess) err = cudaMemcpy(host, dev, sizeof(double) * n, cudaMemcpyDeviceToHost);
        return err == cudaSuccess;
    }
};

int finish(Pool& p, cudaError_t launch, const char* what)
{
    cudaError_t e = (p.err != cudaSuccess) ? p.err : launch;
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        char buf[256];
        std::snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
        return plbm_set_error(buf);
    }
    return 0;
}

PhaseUnits units_of(const plbm_config* c)
{
    PhaseUnits u;
    u.cs2 = c->cs2; u.Kb = c->Kb;
    for (int k = 0; k < 3; ++k) { u.q[k] = c->q[k]; u.m[k] = c->m[k]; }
    return u;
}

// reference parameter order  e, i, n, e_i, e_n, i_n, i_e, n_e, n_i  ->  [species][self, cross0, cross1]
constexpr int EQ_SLOT[3][3] = { { 0, 3, 4 }, { 1, 6, 5 }, { 2, 7, 8 } };

bool have_all(const double* const* p, int n) { for (int i = 0; i < n; ++i) if (!p || !p[i]) return false; return true; }

// every fast division primitive on arbitrary operands, with its validity record
__global__ void selftest_division_kernel(const double* a, const double* b, int n, int mode, double inv_hi, double inv_lo,
                                         double recip_y, double* out, int* ok)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    FastDiv dv;
    D q;
    const double inv[2] = { inv_hi, inv_lo };
    Recip rc; rc.d = b[0]; rc.y = recip_y;
    if (mode == 0) q = dv.xdiv(D(a[i]), D(b[i]));
    else if (mode == 1) q = dv.idiv(D(a[i]), rc, inv);
    else q = dv.cdiv(D(a[i]), rc);
    dv.note_output(q);
    out[i] = q.v;
    ok[i] = dv.ok() ? 1 : 0;
}
__global__ void selftest_recip_kernel(const double* d, double* y) { y[0] = recip_refined(d[0]); }

} // namespace

extern "C" {

int plbm_selftest_division(const double* a, const double* b, int n, int mode, double* out, int* ok)
{
    if (!a || !b || !out || !ok || n < 1) return plbm_set_error("plbm_selftest_division: bad argument");
    Pool p;
    double* da = p.up(a, n); double* db = p.up(b, mode == 0 ? n : 1); double* dout = p.alloc(n);
    int* dok = nullptr;
    if (p.err == cudaSuccess) p.err = cudaMalloc((void**)&dok, sizeof(int) * n);
    if (dok) p.bufs.push_back(dok);
    double* dy = p.alloc(1);
    double y = 0.0, hi = 0.0, lo = 0.0;
    if (mode != 0 && p.err == cudaSuccess) {
        selftest_recip_kernel<<<1, 1>>>(db, dy);
        p.down(&y, dy, 1);
        hi = 1.0 / b[0];
        lo = std::fma(-hi, b[0], 1.0) / b[0];
    }
    cudaError_t le = p.err;
    if (le == cudaSuccess) {
        selftest_division_kernel<<<(n + 255) / 256, 256>>>(da, db, n, mode, hi, lo, y, dout, dok);
        le = cudaGetLastError();
    }
    if (finish(p, le, "plbm_selftest_division")) return 1;
    p.down(out, dout, n);
    if (p.err == cudaSuccess) p.err = cudaMemcpy(ok, dok, sizeof(int) * n, cudaMemcpyDeviceToHost);
    return finish(p, cudaSuccess, "plbm_selftest_division");
}

int plbm_host_update_macro(int NX, int NY, const plbm_config* units, const double* const f[3], const double* const g[3],
                           const double* Ex, const double* Ey, double* const rho[3], double* const ux[3], double* const uy[3],
                           double* const T[3], double* const upx[3], double* const upy[3], double* rho_q)
{
    if (!units || !have_all(f, 3) || !have_all(g, 3) || !Ex || !Ey) return plbm_set_error("plbm_host_update_macro: null argument");
    const size_t N = (size_t)NX * NY;
    Pool p; PhaseArrays a = {};
    for (int k = 0; k < 3; ++k) {
        a.f[k] = p.up(f[k], N * NQ); a.g[k] = p.up(g[k], N * NQ);
        a.rho[k] = p.alloc(N); a.ux[k] = p.alloc(N); a.uy[k] = p.alloc(N); a.T[k] = p.alloc(N); a.upx[k] = p.alloc(N); a.upy[k] = p.alloc(N);
    }
    a.Ex = p.up(Ex, N); a.Ey = p.up(Ey, N); a.rho_q = p.alloc(N);
    cudaError_t le = p.err == cudaSuccess ? launch_update_macro(a, units_of(units), (int)N, 0) : p.err;
    if (finish(p, le, "plbm_host_update_macro")) return 1;
    for (int k = 0; k < 3; ++k) {
        if (rho && rho[k]) p.down(rho[k], a.rho[k], N);
        if (ux && ux[k]) p.down(ux[k], a.ux[k], N);
        if (uy && uy[k]) p.down(uy[k], a.uy[k], N);
        if (T && T[k]) p.down(T[k], a.T[k], N);
        if (upx && upx[k]) p.down(upx[k], a.upx[k], N);
        if (upy && upy[k]) p.down(upy[k], a.upy[k], N);
    }
    if (rho_q) p.down(rho_q, a.rho_q, N);
    return finish(p, cudaSuccess, "plbm_host_update_macro");
}

int plbm_host_equilibrium(int NX, int NY, const plbm_config* units, const double* const rho[3], const double* const ux[3],
                          const double* const uy[3], const double* const T[3], const double* const upx[3], const double* const upy[3],
                          double* const f_eq[9], double* const g_eq[9])
{
    if (!units || !have_all(rho, 3) || !have_all(ux, 3) || !have_all(uy, 3) || !have_all(T, 3) || !have_all(upx, 3) || !have_all(upy, 3) ||
        !have_all(f_eq, 9) || !have_all(g_eq, 9)) return plbm_set_error("plbm_host_equilibrium: null argument");
    const size_t N = (size_t)NX * NY;
    Pool p; PhaseArrays a = {};
    for (int k = 0; k < 3; ++k) {
        a.rho[k] = p.up(rho[k], N); a.ux[k] = p.up(ux[k], N); a.uy[k] = p.up(uy[k], N); a.T[k] = p.up(T[k], N);
        a.upx[k] = p.up(upx[k], N); a.upy[k] = p.up(upy[k], N);
        for (int m = 0; m < 3; ++m) { a.feq[k][m] = p.alloc(N * NQ); a.geq[k][m] = p.alloc(N * NQ); }
    }
    cudaError_t le = p.err == cudaSuccess ? launch_equilibrium(a, units_of(units), (int)N, 0) : p.err;
    if (finish(p, le, "plbm_host_equilibrium")) return 1;
    for (int k = 0; k < 3; ++k)
        for (int m = 0; m < 3; ++m) {
            p.down(f_eq[EQ_SLOT[k][m]], a.feq[k][m], N * NQ);
            p.down(g_eq[EQ_SLOT[k][m]], a.geq[k][m], N * NQ);
        }
    return finish(p, cudaSuccess, "plbm_host_equilibrium");
}

int plbm_host_thermal_collisions(int NX, int NY, const plbm_config* units, const double* const g[3], const double* const g_eq[9],
                                 const double* const f_eq[9], const double* const rho[3], const double* const ux[3],
                                 const double* const uy[3], double* const out[3])
{
    if (!units || !have_all(g, 3) || !have_all(g_eq, 9) || !have_all(f_eq, 9) || !have_all(rho, 3) || !have_all(ux, 3) || !have_all(uy, 3) ||
        !have_all(out, 3)) return plbm_set_error("plbm_host_thermal_collisions: null argument");
    const size_t N = (size_t)NX * NY;
    Pool p; PhaseArrays a = {};
    for (int k = 0; k < 3; ++k) {
        a.g[k] = p.up(g[k], N * NQ); a.tmp[k] = p.alloc(N * NQ);
        a.rho[k] = p.up(rho[k], N); a.ux[k] = p.up(ux[k], N); a.uy[k] = p.up(uy[k], N);
        for (int m = 0; m < 3; ++m) { a.feq[k][m] = p.up(f_eq[EQ_SLOT[k][m]], N * NQ); a.geq[k][m] = p.up(g_eq[EQ_SLOT[k][m]], N * NQ); }
    }
    cudaError_t le = p.err == cudaSuccess ? launch_thermal_collisions(a, units_of(units), (int)N, 0) : p.err;
    if (finish(p, le, "plbm_host_thermal_collisions")) return 1;
    for (int k = 0; k < 3; ++k) p.down(out[k], a.tmp[k], N * NQ);
    return finish(p, cudaSuccess, "plbm_host_thermal_collisions");
}

int plbm_host_collisions(int NX, int NY, const plbm_config* units, const double* const f[3], const double* const f_eq[9],
                         const double* const rho[3], const double* const ux[3], const double* const uy[3],
                         const double* Ex, const double* Ey, double* const out[3])
{
    if (!units || !have_all(f, 3) || !have_all(f_eq, 9) || !have_all(rho, 2) || !have_all(ux, 2) || !have_all(uy, 2) || !Ex || !Ey ||
        !have_all(out, 3)) return plbm_set_error("plbm_host_collisions: null argument");
    const size_t N = (size_t)NX * NY;
    Pool p; PhaseArrays a = {};
    for (int k = 0; k < 3; ++k) {
        a.f[k] = p.up(f[k], N * NQ); a.tmp[k] = p.alloc(N * NQ);
        if (k < 2) { a.rho[k] = p.up(rho[k], N); a.ux[k] = p.up(ux[k], N); a.uy[k] = p.up(uy[k], N); }
        for (int m = 0; m < 3; ++m) a.feq[k][m] = p.up(f_eq[EQ_SLOT[k][m]], N * NQ);
    }
    a.Ex = p.up(Ex, N); a.Ey = p.up(Ey, N);
    cudaError_t le = p.err == cudaSuccess ? launch_collisions(a, units_of(units), (int)N, 0) : p.err;
    if (finish(p, le, "plbm_host_collisions")) return 1;
    for (int k = 0; k < 3; ++k) p.down(out[k], a.tmp[k], N * NQ);
    return finish(p, cudaSuccess, "plbm_host_collisions");
}

int plbm_host_stream(int NX, int NY, int bc_type, const double* const in[3], double* const out[3])
{
    if (!have_all(in, 3) || !have_all(out, 3)) return plbm_set_error("plbm_host_stream: null argument");
    if (bc_type != PLBM_BC_PERIODIC && bc_type != PLBM_BC_BOUNCEBACK) return plbm_set_error("plbm_host_stream: unknown boundary type");
    const size_t n = (size_t)NX * NY * NQ;
    Pool p;
    const double* src[3]; double* dst[3];
    // out[] is in-out: with walls, slots no source reaches keep what the caller's temp_* array held (reference quirk)
    for (int k = 0; k < 3; ++k) { src[k] = p.up(in[k], n); dst[k] = p.up(out[k], n); }
    cudaError_t le = p.err != cudaSuccess ? p.err
                   : (bc_type == PLBM_BC_PERIODIC ? launch_stream_periodic(src, dst, NX, NY, 0) : launch_stream_bounceback(src, dst, NX, NY, 0));
    if (finish(p, le, "plbm_host_stream")) return 1;
    for (int k = 0; k < 3; ++k) p.down(out[k], dst[k], n);
    return finish(p, cudaSuccess, "plbm_host_stream");
}

} // extern "C"
