// plbm_api.cu -- the C ABI of include/plbm.h: context, device state, and the time-step driver.
//
// Host-side mirror of LBmethod's constructor and loop (reference src/plasma.cpp:22-124, 459-529):
// the context owns all state in HBM; a step is K1 (fused pull-stream/moments/equilibria/collisions)
// followed by the Poisson solve of the configured type, exactly in the reference's order.
#include "../../include/plbm.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "exact_math.cuh"
#include "host_tables.h"
#include "charge_pull.h"
#include "k1_fused.h"
#include "layout.h"
#include "lbm_consts.h"
#include "output.h"
#include "phases.h"
#include "poisson_fft.h"
#include "poisson_iter.h"

using namespace plbm;

constexpr int PEER_NBUF = 6;              // buffers a slab exposes to its peers: T1, flags, halo_recv_lo, halo_recv_hi, phi_below, phi_above

namespace {

thread_local std::string g_err;

int fail(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

__global__ void recip_kernel(const double* d, double* y, int n)
{
    const int t = threadIdx.x;
    if (t < n) y[t] = recip_refined(d[t]);
}

// Barrier among the slabs' streams: every rank publishes the epoch in its slot of every peer's flag array and waits
// until all slots of its own array have reached it.  Work enqueued before the barrier on any rank (its peer-memory
// writes included: release at system scope after the preceding kernels completed) is visible to work enqueued after
// it on every rank.  One GPU per rank, so the spinning CTA never keeps a peer's kernels from running.
struct PeerFlags {
    unsigned long long* f[PLBM_MAX_RANKS];
};
__global__ void peer_barrier_kernel(const __grid_constant__ PeerFlags pf, int rank, int nranks, unsigned long long epoch, int* timeout)
{
    const int t = threadIdx.x;
    if (t < nranks) {
        __threadfence_system();
        unsigned long long* theirs = pf.f[t] + rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(theirs), "l"(epoch) : "memory");
        const unsigned long long* mine = pf.f[rank] + t;
        const long long t0 = clock64();
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
            if (v >= epoch) break;
            if (clock64() - t0 > 20000000000ll) { *timeout = 1; break; }     // ~10 s: a peer died; do not hang the GPU
        }
        __threadfence_system();
    }
}

struct PeerBlob {                       // what one rank tells the others (plbm_peer_export)
    cudaIpcMemHandle_t handle[PEER_NBUF];
    long long offset[PEER_NBUF];        // of the pointer inside the IPC-mapped allocation
    int rank, nyl;
};
static_assert(sizeof(PeerBlob) <= PLBM_PEER_BLOB_BYTES, "PLBM_PEER_BLOB_BYTES too small");

// offset of p inside the allocation an IPC handle of p maps (cudaMalloc may sub-allocate)
int allocation_offset(const void* p, long long* off)
{
    typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult st;
    CUDA_TRY(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &st));
    if (!fn || st != cudaDriverEntryPointSuccess) return fail("cuMemGetAddressRange is not available");
    unsigned long long base = 0;
    size_t size = 0;
    if (((range_fn)fn)(&base, &size, (unsigned long long)p) != 0) return fail("cuMemGetAddressRange failed");
    *off = (long long)((unsigned long long)p - base);
    return 0;
}

} // namespace

int plbm_set_error(const char* msg) { return fail("%s", msg); }

struct plbm_ctx {
    plbm_config cfg;
    int device = -1;                         // ordinal every launch of this context targets (entry points switch to it and back)
    LbmConsts consts;
    LbmGeom geom;
    cudaStream_t stream = nullptr;
    double* pop[2] = { nullptr, nullptr };   // ping-pong population planes
    alignas(64) CUtensorMap pop_map[2];      // the same planes as TMA tensors (periodic lattices: the pull of K1 is done by the TMA engine)
    bool tma = false;                        // k1_pool_kernel (PLBM_K1_POOL=1)
    bool tma_tile = false;                   // k1_tma_kernel (PLBM_K1_TMA=1)
    int cur = 0;                             // pop[cur] holds the current post-collision state
    double* Ex = nullptr; double* Ey = nullptr; double* rho_q = nullptr; double* phi = nullptr;
    double* macro[12] = {};                  // ux,uy (e,i,n), T (e,i,n), rho (e,i,n) in plbm.h field order (the set last written)
    double* macro_sets[2][12] = {};          // fused path: alternating sets so that a pending fetch keeps its data
    int macro_cur = 0;
    double* snap[4] = {};                    // rho_q, Ex, Ey, phi as of the last plbm_fetch_begin
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_snap = nullptr, ev_fetched = nullptr;
    bool fetch_pending = false;
    float* frames = nullptr;                 // alternate output path: NUM_FRAMES planes of floats
    double* series = nullptr;                // [NUM_SERIES][NUM_POINTS]
    double* staging = nullptr;               // 9*NX*NYl doubles for AoS transfers
    bool macro_valid = false;
    bool halo_fresh = false;                 // slabs: the halo rows were filled by plbm_initialize / plbm_upload_state themselves (each rank
                                             // writes every row its own cells pull from), so there is nothing to exchange until a step has run
    bool e_stale = false;                    // Ex/Ey arrays not materialised: K1 takes E = -grad(phi) from phi itself (fused periodic FFT path)
    bool poisson_called = false;             // call_once of reference src/poisson.cpp:34-41
    // spectral Poisson
    PoissonFftDev fft = {};
    cpx* tw_row = nullptr; cpx* tw_col = nullptr; double* sx2 = nullptr; double* sy2 = nullptr;
    // multi-slab exchange buffers
    double* halo_send_lo = nullptr; double* halo_send_hi = nullptr; double* halo_recv_lo = nullptr; double* halo_recv_hi = nullptr;
    double* phi_below = nullptr; double* phi_above = nullptr;
    // peer-memory transposes: every slab's T1 and barrier flags mapped through CUDA IPC
    bool peers = false;
    PeerTable peer_t1 = {};
    unsigned long long* flags = nullptr;                 // [PLBM_MAX_RANKS] arrival counters written by the peers
    unsigned long long* peer_flags[PLBM_MAX_RANKS] = {};
    double* up_halo_recv_lo = nullptr;                   // upper neighbour's buffer for what arrives from below (= from me)
    double* down_halo_recv_hi = nullptr;                 // lower neighbour's buffer for what arrives from above
    double* up_phi_below = nullptr;                      // upper neighbour's copy of the row below its slab (= my last row)
    double* down_phi_above = nullptr;
    // Pipelined peer step (plbm_step_peer): the Poisson solve of step t runs on its own stream beside K1 of step t.  It needs a
    // second potential (K1 reads phi of step t-1 while P3 writes phi of step t), second copies of the neighbours' boundary rows,
    // two sets of halo receive buffers (one barrier per exchange instead of three), and its own barrier flags and epoch.
    cudaStream_t pstream = nullptr;
    cudaEvent_t ev_pop = nullptr, ev_phi = nullptr;
    double* rho_q_next = nullptr;                        // charge density pulled from the planes the last step wrote (charge_pull.cu)
    double* phi_buf[2] = { nullptr, nullptr };           // c->phi == phi_buf[phi_par]
    double* phi_below_base = nullptr; double* phi_above_base = nullptr;         // [2][NX]; c->phi_below == base + phi_par*NX
    double* up_phi_below_base = nullptr; double* down_phi_above_base = nullptr; // the neighbours' pairs
    int phi_par = 0;                                     // which potential is the current one (all slabs flip together)
    double mass[2] = { 1.0, 1.0 };                        // m_e, m_i (charge_pull.cu)
    int halo_par = 0;                                    // which halo receive buffers the next exchange uses
    unsigned long long epoch_b = 0;                      // barriers of the Poisson stream (second half of the flag arrays)
    void* peer_mapped[PEER_NBUF * PLBM_MAX_RANKS] = {};  // bases returned by cudaIpcOpenMemHandle
    int* peer_timeout = nullptr;                         // device flag: a barrier gave up waiting
    unsigned long long epoch = 0;
    int slab_y0[PLBM_MAX_RANKS + 1] = {};
    int slab_k0[PLBM_MAX_RANKS + 1] = {};
    // unfused device path (bounce-back walls): the reference's own array set in its own layout
    bool unfused = false;
    // fused walls path (default for bounce-back): K1 with the walls pull (walls.cuh)
    bool walls = false;
    double* rim[2] = { nullptr, nullptr };   // pre-collision f of the wall cells, ping-pong
    int rim_cur = 0;
    bool identity_pull = true;               // pop[cur] is the state at the top of the loop itself
    PhaseArrays pa = {};
    PhaseUnits pu = {};
    std::vector<double*> uf_bufs;
    unsigned long long* err_bits = nullptr;  // iterative solvers: ping-pong max-update slots
    int* iters_dev = nullptr;
    long long bytes = 0;
    std::vector<cudaEvent_t> events;
};

namespace {

// Every extern "C" entry makes the context's device current for its own duration and restores the caller's: two contexts on
// different GPUs can live in one process (plbm_group), and a caller that never selected the device still launches correctly.
struct DevGuard {
    int prev = -1;
    bool switched = false;
    explicit DevGuard(const plbm_ctx* c)
    {
        if (c && c->device >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != c->device)
            switched = (cudaSetDevice(c->device) == cudaSuccess);
    }
    ~DevGuard() { if (switched) cudaSetDevice(prev); }
    DevGuard(const DevGuard&) = delete;
    DevGuard& operator=(const DevGuard&) = delete;
};

// Makes potential `par` (and the matching copies of the neighbours' boundary rows, here and in the neighbours) the current one.
void set_phi_parity(plbm_ctx* c, int par)
{
    c->phi_par = par;
    if (c->phi_buf[par]) c->phi = c->phi_buf[par];
    const size_t NX = (size_t)c->cfg.NX;
    if (c->phi_below_base) { c->phi_below = c->phi_below_base + par * NX; c->phi_above = c->phi_above_base + par * NX; }
    if (c->up_phi_below_base) c->up_phi_below = c->up_phi_below_base + par * NX;
    if (c->down_phi_above_base) c->down_phi_above = c->down_phi_above_base + par * NX;
}

template <class T>
int dev_alloc(plbm_ctx* c, T** p, size_t count)
{
    CUDA_TRY(cudaMalloc((void**)p, sizeof(T) * count));
    c->bytes += (long long)(sizeof(T) * count);
    return 0;
}

int build_consts(plbm_ctx* c)
{
    const plbm_config& cfg = c->cfg;
    LbmConsts& k = c->consts;
    k.invcs2 = 1.0 / cfg.cs2;                              // reference src/plasma.cpp:164
    k.hinvcs2 = 0.5 * k.invcs2;
    k.w[0] = 4.0 / 9.0; k.w[1] = 1.0 / 9.0; k.w[2] = 1.0 / 36.0;   // reference src/plasma.cpp:12-16
    for (int s = 0; s < 2; ++s) {
        k.hq[s] = 0.5 * cfg.q[s];                          // reference src/plasma.cpp:389,409
        k.q[s] = cfg.q[s];
        for (int wc = 0; wc < 3; ++wc) k.wq[s][wc] = k.w[wc] * cfg.q[s];   // reference src/collisions.cpp:154,159
        k.gfac[s] = 1.0 - 1.0 / (2 * (double)TAU_VALUE[s]);
    }
    for (int slot = 0; slot < 6; ++slot) {
        const double tau = (double)TAU_VALUE[slot];
        k.a[slot] = 1.0 - 1.0 / tau;                       // reference src/collisions.cpp:86-96
        k.a2[slot] = 2.0 * k.a[slot];
        k.a4[slot] = 2.0 * k.a2[slot];
    }
    {   // 1/tau = hi + lo: hi = RN(1/tau); 1 - hi*tau is exact in one fma, lo = RN((1 - hi*tau)/tau)
        const double taus[3] = { 3.0, 5.0, 6.0 };
        double* outs[3] = { k.inv3, k.inv5, k.inv6 };
        for (int i = 0; i < 3; ++i) {
            const double hi = 1.0 / taus[i];
            outs[i][0] = hi;
            outs[i][1] = std::fma(-hi, taus[i], 1.0) / taus[i];
        }
    }
    // refined reciprocals of the loop-invariant divisors, produced by the device itself
    const double div_h[7] = { cfg.cs2, cfg.Kb, 3.0, 5.0, 6.0, cfg.m[0], cfg.m[1] };
    double y_h[7];
    double *d_d = nullptr, *d_y = nullptr;
    CUDA_TRY(cudaMalloc(&d_d, sizeof(div_h)));
    CUDA_TRY(cudaMalloc(&d_y, sizeof(y_h)));
    CUDA_TRY(cudaMemcpyAsync(d_d, div_h, sizeof(div_h), cudaMemcpyHostToDevice, c->stream));
    recip_kernel<<<1, 32, 0, c->stream>>>(d_d, d_y, 7);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(y_h, d_y, sizeof(y_h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(d_d); cudaFree(d_y);
    Recip* slots[7] = { &k.cs2, &k.Kb, &k.tau3, &k.tau5, &k.tau6, &k.m[0], &k.m[1] };
    for (int i = 0; i < 7; ++i) {
        // FastDiv's domain proof (exact_math.cuh) assumes loop-invariant divisors within 2^+-40
        if (!std::isfinite(y_h[i]) || !(std::fabs(div_h[i]) >= 0x1p-40) || !(std::fabs(div_h[i]) <= 0x1p40))
            return fail("divisor %d (%g) is outside the range the fast division supports", i, div_h[i]);
        slots[i]->d = div_h[i];
        slots[i]->y = y_h[i];
    }
    return 0;
}

int build_fft(plbm_ctx* c)
{
    const int NX = c->cfg.NX, NY = c->cfg.NY, R = c->cfg.nranks;
    if (NX > FFT_MAX_N || NY > FFT_MAX_N) return fail("spectral Poisson supports NX, NY <= %d (got %dx%d)", FFT_MAX_N, NX, NY);
    if (R > 1 && NX != NY) return fail("spectral Poisson on several slabs needs NX == NY (the reference reshapes the flat array as NX rows of NY, src/poisson.cpp:621)");
    const int n0 = NX, n1 = NY, nh = n1 / 2 + 1, nyl = (R > 1) ? c->geom.NYl : n0;
    std::vector<double> tw((size_t)2 * (n0 > n1 ? n0 : n1));
    if (dev_alloc(c, &c->tw_row, n1)) return 1;
    host_twiddles(n1, tw.data());
    CUDA_TRY(cudaMemcpy(c->tw_row, tw.data(), sizeof(cpx) * n1, cudaMemcpyHostToDevice));
    if (dev_alloc(c, &c->tw_col, n0)) return 1;
    host_twiddles(n0, tw.data());
    CUDA_TRY(cudaMemcpy(c->tw_col, tw.data(), sizeof(cpx) * n0, cudaMemcpyHostToDevice));
    std::vector<double> s2((size_t)(n0 > nh ? n0 : nh));
    if (dev_alloc(c, &c->sx2, n0)) return 1;
    host_sin2_rows(NX, s2.data());
    CUDA_TRY(cudaMemcpy(c->sx2, s2.data(), sizeof(double) * n0, cudaMemcpyHostToDevice));
    if (dev_alloc(c, &c->sy2, nh)) return 1;
    host_sin2_cols(NY, s2.data());
    CUDA_TRY(cudaMemcpy(c->sy2, s2.data(), sizeof(double) * nh, cudaMemcpyHostToDevice));
    PoissonFftDev& f = c->fft;
    f.n0 = n0; f.n1 = n1; f.nyl = nyl;
    f.tab.nranks = R;
    for (int r = 0; r <= R; ++r) {
        f.tab.y0[r] = (R > 1) ? c->slab_y0[r] : (r == 0 ? 0 : n0);
        c->slab_k0[r] = (int)(((long long)nh * r) / R);
    }
    f.k0 = c->slab_k0[c->cfg.rank];
    f.tab.nkl = c->slab_k0[c->cfg.rank + 1] - f.k0;
    if (dev_alloc(c, &f.T1, (size_t)nh * nyl)) return 1;
    if (R > 1) {
        if (dev_alloc(c, &f.T2, (size_t)(f.tab.nkl > 0 ? f.tab.nkl : 1) * n0)) return 1;
        if (dev_alloc(c, &f.p2_flags, (size_t)f.tab.nkl + 2)) return 1;            // poisson_cols_gather_kernel
        CUDA_TRY(cudaMemset(f.p2_flags, 0, sizeof(unsigned) * ((size_t)f.tab.nkl + 2)));
        f.p2_epoch = 0;
        f.p2_per_sm = 0;
    }
    else f.T2 = f.T1;
    f.row = make_fft_plan(n1, c->tw_row);
    f.col = make_fft_plan(n0, c->tw_col);
    f.sx2 = c->sx2; f.sy2 = c->sy2;
    f.norm = 1.0 / (NX * NY);                              // reference src/poisson.cpp:415
    CUDA_TRY(poisson_fft_configure(f));
    return 0;
}

// call_once part of poisson::SolvePoisson, reference src/poisson.cpp:34-41
int poisson_first_call(plbm_ctx* c)
{
    if (c->poisson_called) return 0;
    c->poisson_called = true;
    const size_t n = (size_t)c->cfg.NX * c->geom.NYl;
    CUDA_TRY(cudaMemsetAsync(c->phi, 0, sizeof(double) * n, c->stream));
    if (c->cfg.poisson_type == PLBM_POISSON_NONE) {
        CUDA_TRY(cudaMemsetAsync(c->Ex, 0, sizeof(double) * n, c->stream));
        CUDA_TRY(cudaMemsetAsync(c->Ey, 0, sizeof(double) * n, c->stream));
    }
    return 0;
}

// one of poisson::SolvePoisson_{GS,SOR,FFT,9point}[_Periodic]: rho_q -> phi
int poisson_solver(plbm_ctx* c, int type, long long* launches)
{
    if (type == PLBM_POISSON_FFT) {
        if (!c->fft.T1 && build_fft(c)) return 1;            // plan and tables on the first spectral request (host API: the type is per call)
        if (c->cfg.nranks != 1) return fail("several slabs: drive the Poisson stages with plbm_poisson_stage() around the all-to-all exchanges");
        CUDA_TRY(launch_poisson_rows_fwd(c->fft, c->rho_q, c->stream));
        CUDA_TRY(launch_poisson_cols(c->fft, c->stream));
        CUDA_TRY(launch_poisson_rows_inv(c->fft, c->phi, c->stream));
        if (launches) *launches += 3;
        return 0;
    }
    const bool wrapped = (type == PLBM_POISSON_GS_PERIODIC || type == PLBM_POISSON_SOR_PERIODIC || type == PLBM_POISSON_NPS_PERIODIC);
    if (type == PLBM_POISSON_GS || type == PLBM_POISSON_SOR || type == PLBM_POISSON_NPS || wrapped) {
        if (c->cfg.nranks != 1) return fail("the iterative Poisson solvers run on a single slab");
        const int kind = (type == PLBM_POISSON_GS || type == PLBM_POISSON_GS_PERIODIC) ? 0
                       : (type == PLBM_POISSON_SOR || type == PLBM_POISSON_SOR_PERIODIC) ? 1 : 2;
        CUDA_TRY(launch_poisson_iterative(kind, wrapped, c->phi, c->rho_q, c->cfg.NX, c->cfg.NY, c->cfg.omega_sor, c->err_bits, c->iters_dev, c->stream));
        if (launches) *launches += 1;
        return 0;
    }
    return fail("unknown Poisson type %d", type);
}

// poisson::ComputeElectricField_Periodic / ComputeElectricField: phi -> Ex, Ey
int poisson_efield(plbm_ctx* c, int bc, long long* launches)
{
    if (bc == PLBM_BC_PERIODIC) {
        const bool slabs = c->cfg.nranks > 1;
        CUDA_TRY(launch_efield_periodic(c->phi, slabs ? c->phi_below : nullptr, slabs ? c->phi_above : nullptr, c->Ex, c->Ey,
                                        c->cfg.NX, c->geom.NYl, c->stream));
        if (launches) *launches += 1;
        return 0;
    }
    if (c->cfg.nranks != 1) return fail("field reconstruction with walls runs on a single slab");
    CUDA_TRY(launch_efield_walls(c->phi, c->Ex, c->Ey, c->cfg.NX, c->cfg.NY, c->stream));
    if (launches) *launches += 3;
    return 0;
}

// poisson::SolvePoisson dispatch, reference src/poisson.cpp:25-82
int solve_poisson(plbm_ctx* c, long long* launches)
{
    const int type = c->cfg.poisson_type, bc = c->cfg.bc_type;
    if (poisson_first_call(c)) return 1;
    if (type == PLBM_POISSON_NONE) return 0;
    if (bc != PLBM_BC_PERIODIC && type == PLBM_POISSON_FFT) return 0;   // reference src/poisson.cpp:76-77
    if (poisson_solver(c, type, launches)) return 1;
    return poisson_efield(c, bc, launches);
}

// Ex/Ey as arrays, for everything that is not K1 (downloads, host API, walls)
int materialise_efield(plbm_ctx* c, long long* launches = nullptr)
{
    if (!c->e_stale) return 0;
    c->e_stale = false;
    return poisson_efield(c, PLBM_BC_PERIODIC, launches);
}

// the time loop's SolvePoisson: on the fused periodic spectral path the field stays implicit in phi
int solve_poisson_in_loop(plbm_ctx* c, long long* launches)
{
    const bool implicit = !c->unfused && c->pop[0] && c->cfg.poisson_type == PLBM_POISSON_FFT && c->cfg.bc_type == PLBM_BC_PERIODIC;
    if (!implicit) return solve_poisson(c, launches);
    if (poisson_first_call(c)) return 1;
    if (poisson_solver(c, PLBM_POISSON_FFT, launches)) return 1;
    c->e_stale = true;
    return 0;
}

int one_step(plbm_ctx* c, bool want_fields, long long* launches)
{
    if (c->unfused) {
        // the reference's own sequence of sweeps (src/plasma.cpp:478-504) on its own array set; temp_* keeps the
        // stale contents the bounce-back streaming lets through (src/streaming.cpp:66-112)
        const int N = c->cfg.NX * c->cfg.NY;
        PhaseArrays& a = c->pa;
        CUDA_TRY(launch_update_macro(a, c->pu, N, c->stream));
        CUDA_TRY(launch_equilibrium(a, c->pu, N, c->stream));
        CUDA_TRY(launch_thermal_collisions(a, c->pu, N, c->stream));
        for (int k = 0; k < 3; ++k) std::swap(a.g[k], a.tmp[k]);
        CUDA_TRY(launch_collisions(a, c->pu, N, c->stream));
        for (int k = 0; k < 3; ++k) std::swap(a.f[k], a.tmp[k]);
        CUDA_TRY(launch_stream_bounceback(a.f, a.tmp, c->cfg.NX, c->cfg.NY, c->stream));
        for (int k = 0; k < 3; ++k) std::swap(a.f[k], a.tmp[k]);
        CUDA_TRY(launch_stream_bounceback(a.g, a.tmp, c->cfg.NX, c->cfg.NY, c->stream));
        for (int k = 0; k < 3; ++k) std::swap(a.g[k], a.tmp[k]);
        if (launches) *launches += 6;
        c->macro_valid = true;
        return 0;
    }
    if (!c->pop[0]) return fail("plbm_step: context was created with fields_only");
    if (want_fields && c->macro_sets[1][0]) {              // fetch pipeline in use: write the set no pending copy reads
        c->macro_cur ^= 1;
        for (int k = 0; k < 12; ++k) c->macro[k] = c->macro_sets[c->macro_cur][k];
    }
    MacroOut mo;
    for (int s = 0; s < 3; ++s) {
        mo.ux[s] = c->macro[2 * s]; mo.uy[s] = c->macro[2 * s + 1];
        mo.T[s] = c->macro[6 + s]; mo.rho[s] = c->macro[9 + s];
    }
    if (c->walls) {
        const WallArgs wa = { c->rim[c->rim_cur], c->rim[c->rim_cur ^ 1], wall_rim_count(c->cfg.NX, c->geom.NYl), c->identity_pull ? 1 : 0 };
        CUDA_TRY(launch_k1_fused_walls(c->pop[c->cur], c->pop[c->cur ^ 1], c->Ex, c->Ey, c->rho_q, want_fields ? &mo : nullptr,
                                       c->consts, c->geom, wa, c->stream));
        c->rim_cur ^= 1;
        c->identity_pull = false;
    } else if (c->tma) {
        const bool slabs = c->cfg.nranks > 1;
        CUDA_TRY(launch_k1_pool(c->pop_map[c->cur], c->pop[c->cur], c->pop[c->cur ^ 1], c->Ex, c->Ey, c->e_stale ? c->phi : nullptr,
                               slabs ? c->phi_below : nullptr, slabs ? c->phi_above : nullptr, c->rho_q, want_fields ? &mo : nullptr,
                               c->consts, c->geom, c->stream));
    } else if (c->tma_tile) {
        const bool slabs = c->cfg.nranks > 1;
        CUDA_TRY(launch_k1_tma(c->pop_map[c->cur], c->pop[c->cur], c->pop[c->cur ^ 1], c->Ex, c->Ey, c->e_stale ? c->phi : nullptr,
                               slabs ? c->phi_below : nullptr, slabs ? c->phi_above : nullptr, c->rho_q, want_fields ? &mo : nullptr,
                               c->consts, c->geom, c->stream));
    } else if (c->e_stale) {
        const bool slabs = c->cfg.nranks > 1;
        CUDA_TRY(launch_k1_fused_phi(c->pop[c->cur], c->pop[c->cur ^ 1], c->phi, slabs ? c->phi_below : nullptr, slabs ? c->phi_above : nullptr,
                                     c->rho_q, want_fields ? &mo : nullptr, c->consts, c->geom, c->stream));
    } else {
        CUDA_TRY(launch_k1_fused(c->pop[c->cur], c->pop[c->cur ^ 1], c->Ex, c->Ey, c->rho_q, want_fields ? &mo : nullptr,
                                 c->consts, c->geom, c->stream));
    }
    if (launches) *launches += 1;
    c->cur ^= 1;
    c->macro_valid = want_fields;
    c->halo_fresh = false;
    return 0;
}

} // namespace

extern "C" {

const char* plbm_last_error(void) { return g_err.c_str(); }

int plbm_units_from_si(int Z_ion, int A_ion, double Ex_SI, double Ey_SI, double T_e_SI, double T_i_SI, double T_n_SI,
                       double n_e_SI, double n_n_SI, plbm_config* o)
{
    if (!o) return fail("plbm_units_from_si: null config");
    // reference include/plasma.hpp:76-133, same expression order
    const double kB_SI = 1.380649e-23, e_charge_SI = 1.602176634e-19, epsilon0_SI = 8.854187817e-12;
    const double m_e_SI = 9.10938356e-31, u_SI = 1.66053906660e-27;
    const double m_i_SI = A_ion * u_SI, m_n_SI = A_ion * u_SI;
    const double n0_SI = n_e_SI, M0_SI = m_e_SI, T0_SI = T_e_SI, Q0_SI = e_charge_SI;
    const double L0_SI = std::sqrt(epsilon0_SI * kB_SI * T0_SI / (n0_SI * Q0_SI * Q0_SI)) * 1e-2;
    const double t0_SI = std::sqrt(epsilon0_SI * M0_SI / (3.0 * n0_SI * Q0_SI * Q0_SI)) * 1e-2;
    const double E0_SI = M0_SI * L0_SI / (Q0_SI * t0_SI * t0_SI);
    o->cs2 = kB_SI * T0_SI / M0_SI * t0_SI * t0_SI / (L0_SI * L0_SI);
    o->Kb = kB_SI * (t0_SI * t0_SI * T0_SI) / (L0_SI * L0_SI * M0_SI);
    o->Ex_ext = Ex_SI / E0_SI;
    o->Ey_ext = Ey_SI / E0_SI;
    o->T_init[0] = T_e_SI / T0_SI; o->T_init[1] = T_i_SI / T0_SI; o->T_init[2] = T_n_SI / T0_SI;
    o->m[0] = m_e_SI / M0_SI; o->m[1] = m_i_SI / M0_SI; o->m[2] = m_n_SI / M0_SI;
    o->q[0] = -e_charge_SI / Q0_SI; o->q[1] = Z_ion * e_charge_SI / Q0_SI; o->q[2] = 0.0;
    o->rho_init[0] = o->m[0] * n_e_SI / n0_SI;
    o->rho_init[1] = o->m[1] * n_e_SI / n0_SI / Z_ion;
    o->rho_init[2] = o->m[2] * n_n_SI / n0_SI;
    return 0;
}

int plbm_create(const plbm_config* cfg, plbm_ctx** out)
{
    if (!cfg || !out) return fail("plbm_create: null argument");
    *out = nullptr;
    if (cfg->NX < 3 || cfg->NY < 3) return fail("plbm_create: lattice %dx%d too small", cfg->NX, cfg->NY);
    if ((long long)cfg->NX * cfg->NY * 9 >= (1LL << 31))
        return fail("plbm_create: NX*NY*9 must stay below 2^31 like the reference's int indices (reference src/plasma.cpp:58)");
    if (cfg->poisson_type < 0 || cfg->poisson_type > 4) return fail("plbm_create: unknown Poisson type %d", cfg->poisson_type);
    if (cfg->bc_type != PLBM_BC_PERIODIC && cfg->bc_type != PLBM_BC_BOUNCEBACK) return fail("plbm_create: unknown boundary type %d", cfg->bc_type);
    if (cfg->nranks > 1 && (cfg->bc_type != PLBM_BC_PERIODIC || (cfg->poisson_type != PLBM_POISSON_FFT && cfg->poisson_type != PLBM_POISSON_NONE)))
        return fail("plbm_create: several slabs support periodic boundaries with the spectral (or no) Poisson solve only");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("plbm_create: no CUDA device (this library has no CPU path)");
    int caller_device = 0;
    CUDA_TRY(cudaGetDevice(&caller_device));
    struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{ caller_device };     // the caller's device is current again on return
    if (cfg->device >= 0) CUDA_TRY(cudaSetDevice(cfg->device));

    plbm_ctx* c = new plbm_ctx();
    c->cfg = *cfg;
    c->device = cfg->device >= 0 ? cfg->device : caller_device;
    if (c->cfg.nranks <= 1) { c->cfg.nranks = 1; c->cfg.rank = 0; c->cfg.y0 = 0; c->cfg.NY_local = cfg->NY; }
    else {
        // the library owns the decomposition rule (plbm_slab_of); y0 / NY_local of the caller are ignored
        if (c->cfg.nranks > PLBM_MAX_RANKS || c->cfg.rank < 0 || c->cfg.rank >= c->cfg.nranks) { delete c; return fail("plbm_create: rank %d of %d", cfg->rank, cfg->nranks); }
        for (int r = 0; r <= c->cfg.nranks; ++r) {
            int y0r = 0, nylr = 0;
            if (r < c->cfg.nranks && plbm_slab_of(cfg->NY, r, c->cfg.nranks, &y0r, &nylr)) { delete c; return 1; }
            c->slab_y0[r] = (r < c->cfg.nranks) ? y0r : cfg->NY;
        }
        c->cfg.y0 = c->slab_y0[c->cfg.rank];
        c->cfg.NY_local = c->slab_y0[c->cfg.rank + 1] - c->cfg.y0;
    }
    c->mass[0] = cfg->m[0]; c->mass[1] = cfg->m[1];
    c->geom.NX = cfg->NX;
    c->geom.NYl = c->cfg.NY_local;
    c->geom.pitch = ((cfg->NX + 15) / 16) * 16;
    c->geom.plane = (long long)c->geom.pitch * (c->geom.NYl + 2);
    c->geom.wrap_y = (c->cfg.nranks == 1) ? 1 : 0;
    c->geom.prefetch_rows = (cfg->bc_type == PLBM_BC_PERIODIC && !cfg->fields_only) ? k1_prefetch_rows(cfg->NX, c->geom.NYl) : 0;
    if (c->geom.plane >= (1LL << 31)) { delete c; return fail("plbm_create: slab too large for 32-bit plane offsets"); }

#define TRY_OR_DESTROY(expr) do { if ((expr) != 0) { std::string keep = g_err; plbm_destroy(c); g_err = keep; return 1; } } while (0)
#define CUDA_OR_DESTROY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { fail("%s failed: %s", #expr, cudaGetErrorString(e__)); std::string keep = g_err; plbm_destroy(c); g_err = keep; return 1; } } while (0)
    CUDA_OR_DESTROY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    const size_t n = (size_t)cfg->NX * c->geom.NYl;
    const size_t pop_count = (size_t)NPLANES * c->geom.plane;
    {
        const char* e = std::getenv("PLBM_UNFUSED_WALLS");       // the reference's own sweep sequence instead of the fused kernel
        const bool want_unfused = e && e[0] && e[0] != '0';
        const bool bb = (cfg->bc_type == PLBM_BC_BOUNCEBACK) && !cfg->fields_only;
        c->unfused = bb && want_unfused;
        c->walls = bb && !want_unfused;
    }
    for (int b = 0; b < 2 && !cfg->fields_only && !c->unfused; ++b) {
        TRY_OR_DESTROY(dev_alloc(c, &c->pop[b], pop_count));
        CUDA_OR_DESTROY(cudaMemsetAsync(c->pop[b], 0, sizeof(double) * pop_count, c->stream));
    }
    if (c->pop[0] && !c->walls) {
        // periodic lattices: PLBM_K1_POOL=1 runs K1 as persistent warps fed by the TMA engine (k1_pool_kernel) instead of the kernel
        // in which every thread pulls its own populations (k1_fused_kernel).  Bit-identical, measured 20 % slower on B200
        // (profiles/r2_k1_sweeps.md), so it is not the default.
        const char* e = std::getenv("PLBM_K1_POOL");
        c->tma = (e && e[0] == '1');
        // default: one CTA per tile, the pull done by the TMA engine (k1_tma_kernel: 0.835 ms at 2048^2 against 0.875 ms for
        // k1_fused_kernel, in which every thread pulls its own populations; PLBM_K1_TMA=0 selects that one)
        const char* e2 = std::getenv("PLBM_K1_TMA");
        c->tma_tile = !c->tma && !(e2 && e2[0] == '0');
        for (int b = 0; b < 2 && (c->tma || c->tma_tile); ++b) CUDA_OR_DESTROY(make_k1_tensor_map(&c->pop_map[b], c->pop[b], c->geom, c->tma));
    }
    TRY_OR_DESTROY(dev_alloc(c, &c->Ex, n));
    TRY_OR_DESTROY(dev_alloc(c, &c->Ey, n));
    TRY_OR_DESTROY(dev_alloc(c, &c->rho_q, n));
    TRY_OR_DESTROY(dev_alloc(c, &c->phi, n));
    for (int k = 0; k < 12 && !cfg->fields_only; ++k) {
        TRY_OR_DESTROY(dev_alloc(c, &c->macro_sets[0][k], n));
        c->macro[k] = c->macro_sets[0][k];
    }
    if (!cfg->fields_only && !c->unfused) TRY_OR_DESTROY(dev_alloc(c, &c->staging, n * NQ));
    if (c->walls) {
        const size_t nr = (size_t)3 * NQ * wall_rim_count(cfg->NX, c->geom.NYl);
        for (int b = 0; b < 2; ++b) {
            TRY_OR_DESTROY(dev_alloc(c, &c->rim[b], nr));
            CUDA_OR_DESTROY(cudaMemsetAsync(c->rim[b], 0, sizeof(double) * nr, c->stream));
        }
    }
    TRY_OR_DESTROY(dev_alloc(c, &c->err_bits, 2));
    TRY_OR_DESTROY(dev_alloc(c, &c->iters_dev, 1));
    if (c->unfused) {
        auto grab = [&](double** p, size_t count) -> int {
            if (dev_alloc(c, p, count)) return 1;
            c->uf_bufs.push_back(*p);
            return cudaMemsetAsync(*p, 0, sizeof(double) * count, c->stream) == cudaSuccess ? 0 : fail("cudaMemsetAsync failed");
        };
        for (int k = 0; k < 3; ++k) {
            TRY_OR_DESTROY(grab(&c->pa.f[k], n * NQ)); TRY_OR_DESTROY(grab(&c->pa.g[k], n * NQ)); TRY_OR_DESTROY(grab(&c->pa.tmp[k], n * NQ));
            for (int m = 0; m < 3; ++m) { TRY_OR_DESTROY(grab(&c->pa.feq[k][m], n * NQ)); TRY_OR_DESTROY(grab(&c->pa.geq[k][m], n * NQ)); }
            TRY_OR_DESTROY(grab(&c->pa.upx[k], n)); TRY_OR_DESTROY(grab(&c->pa.upy[k], n));
            c->pa.ux[k] = c->macro[2 * k]; c->pa.uy[k] = c->macro[2 * k + 1]; c->pa.T[k] = c->macro[6 + k]; c->pa.rho[k] = c->macro[9 + k];
        }
        c->pa.Ex = c->Ex; c->pa.Ey = c->Ey; c->pa.rho_q = c->rho_q;
        c->pu.cs2 = cfg->cs2; c->pu.Kb = cfg->Kb;
        for (int k = 0; k < 3; ++k) { c->pu.q[k] = cfg->q[k]; c->pu.m[k] = cfg->m[k]; }
    }
    CUDA_OR_DESTROY(cudaMemsetAsync(c->rho_q, 0, sizeof(double) * n, c->stream));
    CUDA_OR_DESTROY(cudaMemsetAsync(c->phi, 0, sizeof(double) * n, c->stream));
    CUDA_OR_DESTROY(launch_fill(c->Ex, cfg->Ex_ext, n, c->stream));        // reference src/plasma.cpp:116-117
    CUDA_OR_DESTROY(launch_fill(c->Ey, cfg->Ey_ext, n, c->stream));
    if (c->cfg.nranks > 1) {
        const size_t hn = (size_t)18 * cfg->NX;
        TRY_OR_DESTROY(dev_alloc(c, &c->halo_send_lo, hn)); TRY_OR_DESTROY(dev_alloc(c, &c->halo_send_hi, hn));
        TRY_OR_DESTROY(dev_alloc(c, &c->halo_recv_lo, 2 * hn)); TRY_OR_DESTROY(dev_alloc(c, &c->halo_recv_hi, 2 * hn));     // two sets: halo_par
        TRY_OR_DESTROY(dev_alloc(c, &c->phi_below_base, (size_t)2 * cfg->NX)); TRY_OR_DESTROY(dev_alloc(c, &c->phi_above_base, (size_t)2 * cfg->NX));
        CUDA_OR_DESTROY(cudaMemsetAsync(c->phi_below_base, 0, sizeof(double) * 2 * cfg->NX, c->stream));
        CUDA_OR_DESTROY(cudaMemsetAsync(c->phi_above_base, 0, sizeof(double) * 2 * cfg->NX, c->stream));
        c->phi_below = c->phi_below_base; c->phi_above = c->phi_above_base;
    }
    c->phi_buf[0] = c->phi;
    TRY_OR_DESTROY(build_consts(c));
    if (cfg->poisson_type == PLBM_POISSON_FFT && cfg->bc_type == PLBM_BC_PERIODIC) TRY_OR_DESTROY(build_fft(c));
    CUDA_OR_DESTROY(cudaStreamSynchronize(c->stream));
#undef TRY_OR_DESTROY
#undef CUDA_OR_DESTROY
    *out = c;
    return 0;
}

void plbm_destroy(plbm_ctx* c)
{
    DevGuard guard__(c);
    if (!c) return;
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto e : c->events) cudaEventDestroy(e);
    for (int b = 0; b < 2; ++b) cudaFree(c->pop[b]);
    cudaFree(c->Ex); cudaFree(c->Ey); cudaFree(c->rho_q); cudaFree(c->phi_buf[0] ? c->phi_buf[0] : c->phi);
    for (void* m : c->peer_mapped) if (m) cudaIpcCloseMemHandle(m);
    cudaFree(c->flags); cudaFree(c->peer_timeout);
    for (int b = 0; b < 2; ++b) for (int k = 0; k < 12; ++k) cudaFree(c->macro_sets[b][k]);
    for (int k = 0; k < 4; ++k) cudaFree(c->snap[k]);
    cudaFree(c->frames); cudaFree(c->series);
    cudaFree(c->rim[0]); cudaFree(c->rim[1]);
    if (c->ev_snap) cudaEventDestroy(c->ev_snap);
    if (c->ev_fetched) cudaEventDestroy(c->ev_fetched);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    cudaFree(c->staging);
    for (double* p : c->uf_bufs) cudaFree(p);
    cudaFree(c->err_bits); cudaFree(c->iters_dev);
    cudaFree(c->tw_row); cudaFree(c->tw_col); cudaFree(c->sx2); cudaFree(c->sy2);
    if (c->fft.T2 != c->fft.T1) cudaFree(c->fft.T2);
    cudaFree(c->fft.p2_flags);
    cudaFree(c->fft.T1);
    cudaFree(c->halo_send_lo); cudaFree(c->halo_send_hi); cudaFree(c->halo_recv_lo); cudaFree(c->halo_recv_hi);
    cudaFree(c->phi_below_base); cudaFree(c->phi_above_base);
    cudaFree(c->phi_buf[1]); cudaFree(c->rho_q_next);
    if (c->pstream) cudaStreamDestroy(c->pstream);
    if (c->ev_pop) cudaEventDestroy(c->ev_pop);
    if (c->ev_phi) cudaEventDestroy(c->ev_phi);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int plbm_initialize(plbm_ctx* c)
{
    DevGuard guard__(c);
    if (!c) return fail("plbm_initialize: null context");
    if (c->unfused) {
        const size_t n = (size_t)c->cfg.NX * c->geom.NYl;
        CUDA_TRY(launch_init_aos(c->pa.f, c->pa.g, c->cfg.NX, c->cfg.NY, c->cfg.rho_init, c->cfg.T_init, c->stream));
        for (int k = 0; k < 3; ++k) CUDA_TRY(cudaMemsetAsync(c->pa.tmp[k], 0, sizeof(double) * n * NQ, c->stream));   // std::vector::resize zero-fills temp_*
        CUDA_TRY(launch_fill(c->Ex, c->cfg.Ex_ext, n, c->stream));
        CUDA_TRY(launch_fill(c->Ey, c->cfg.Ey_ext, n, c->stream));
        c->e_stale = false;
        CUDA_TRY(cudaMemsetAsync(c->phi, 0, sizeof(double) * n, c->stream));
        CUDA_TRY(cudaMemsetAsync(c->rho_q, 0, sizeof(double) * n, c->stream));
        c->poisson_called = false;
        c->macro_valid = false;
        return 0;
    }
    if (!c->pop[0]) return fail("plbm_initialize: context was created with fields_only");
    if (c->pstream) CUDA_TRY(cudaStreamSynchronize(c->pstream));
    set_phi_parity(c, 0);          // every slab restarts with the same potential and halo buffers current
    c->halo_par = 0;
    // the state of a freshly constructed LBmethod (reference src/plasma.cpp:58-124): initial populations,
    // E = E_ext, phi = 0 and the Poisson module's call_once not yet taken
    const size_t n = (size_t)c->cfg.NX * c->geom.NYl;
    if (c->walls) {
        CUDA_TRY(launch_initialize_identity(c->pop[c->cur], c->geom, c->cfg.NY, c->cfg.rho_init, c->cfg.T_init, c->consts.w, c->stream));
        c->identity_pull = true;
    } else {
        CUDA_TRY(launch_initialize(c->pop[c->cur], c->geom, c->cfg.NY, c->cfg.y0, c->cfg.rho_init, c->cfg.T_init, c->consts.w, c->stream));
    }
    CUDA_TRY(launch_fill(c->Ex, c->cfg.Ex_ext, n, c->stream));
    CUDA_TRY(launch_fill(c->Ey, c->cfg.Ey_ext, n, c->stream));
    c->e_stale = false;
    CUDA_TRY(cudaMemsetAsync(c->phi, 0, sizeof(double) * n, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->rho_q, 0, sizeof(double) * n, c->stream));
    c->poisson_called = false;
    c->macro_valid = false;
    c->halo_fresh = true;
    return 0;
}

int plbm_upload_state(plbm_ctx* c, const double* const f[3], const double* const g[3])
{
    DevGuard guard__(c);
    if (!c || !f || !g) return fail("plbm_upload_state: null argument");
    if (c->unfused) {
        const size_t ub = sizeof(double) * (size_t)c->geom.NX * c->geom.NYl * NQ;
        for (int s = 0; s < 3; ++s) {
            if (!f[s] || !g[s]) return fail("plbm_upload_state: null array (species %d)", s);
            CUDA_TRY(cudaMemcpyAsync(c->pa.f[s], f[s], ub, cudaMemcpyHostToDevice, c->stream));
            CUDA_TRY(cudaMemcpyAsync(c->pa.g[s], g[s], ub, cudaMemcpyHostToDevice, c->stream));
        }
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        return 0;
    }
    if (!c->pop[0]) return fail("plbm_upload_state: context was created with fields_only");
    const size_t bytes = sizeof(double) * (size_t)c->geom.NX * c->geom.NYl * NQ;
    for (int s = 0; s < 3; ++s)
        for (int kind = 0; kind < 2; ++kind) {
            const double* h = kind ? g[s] : f[s];
            if (!h) return fail("plbm_upload_state: null array (species %d)", s);
            CUDA_TRY(cudaMemcpyAsync(c->staging, h, bytes, cudaMemcpyHostToDevice, c->stream));
            if (c->walls) {
                CUDA_TRY(launch_aos_to_soa_identity(c->staging, c->pop[c->cur], s, kind, c->geom, c->stream));
            } else {
                CUDA_TRY(launch_aos_to_soa(c->staging, c->pop[c->cur], s, kind, c->geom, c->stream));
            }
        }
    if (c->walls) c->identity_pull = true;
    // On a slab the upload has also written the two halo rows (the cells of rows 0 and NYl-1 pull from them), while the rows that
    // LEAVE the slab (directions 2/5/6 of storage row NYl, 4/7/8 of row 1) hold nothing new: an exchange now would overwrite the
    // neighbours' correct halo rows with stale data, so plbm_halo_pack/push/unpack are no-ops until a step has run.
    c->halo_fresh = true;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

int plbm_download_state(plbm_ctx* c, double* const f[3], double* const g[3])
{
    DevGuard guard__(c);
    if (!c || !f || !g) return fail("plbm_download_state: null argument");
    if (c->unfused) {
        const size_t ub = sizeof(double) * (size_t)c->geom.NX * c->geom.NYl * NQ;
        for (int s = 0; s < 3; ++s) {
            if (f[s]) CUDA_TRY(cudaMemcpyAsync(f[s], c->pa.f[s], ub, cudaMemcpyDeviceToHost, c->stream));
            if (g[s]) CUDA_TRY(cudaMemcpyAsync(g[s], c->pa.g[s], ub, cudaMemcpyDeviceToHost, c->stream));
        }
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        return 0;
    }
    if (!c->pop[0]) return fail("plbm_download_state: context was created with fields_only");
    const size_t bytes = sizeof(double) * (size_t)c->geom.NX * c->geom.NYl * NQ;
    for (int s = 0; s < 3; ++s)
        for (int kind = 0; kind < 2; ++kind) {
            double* h = kind ? g[s] : f[s];
            if (!h) continue;
            if (c->walls) {
                CUDA_TRY(launch_soa_to_aos_walls(c->pop[c->cur], c->rim[c->rim_cur], wall_rim_count(c->cfg.NX, c->geom.NYl), c->staging, s, kind,
                                                 c->geom, c->identity_pull ? 1 : 0, c->stream));
            } else {
                CUDA_TRY(launch_soa_to_aos(c->pop[c->cur], c->staging, s, kind, c->geom, c->stream));
            }
            CUDA_TRY(cudaMemcpyAsync(h, c->staging, bytes, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
        }
    return 0;
}

int plbm_set_efield(plbm_ctx* c, const double* Ex, const double* Ey)
{
    DevGuard guard__(c);
    if (!c || !Ex || !Ey) return fail("plbm_set_efield: null argument");
    const size_t bytes = sizeof(double) * (size_t)c->geom.NX * c->geom.NYl;
    c->e_stale = false;
    CUDA_TRY(cudaMemcpyAsync(c->Ex, Ex, bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->Ey, Ey, bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// Small lattices are launch-bound (at 200 x 200 every kernel is less than one wave and a step is four launches): long
// runs of steps are replayed from a CUDA graph of TWO steps (two, so that the ping-pong buffers are back where they were).
// The first two steps run eagerly -- they settle everything a later step does not repeat (call_once of the Poisson module,
// field source, kernel attributes) -- and so does the tail, including the step whose moments are asked for.
constexpr long long GRAPH_MAX_CELLS = 1 << 20;
constexpr int GRAPH_MIN_STEPS = 24;

int plbm_step(plbm_ctx* c, int nsteps, int want_fields)
{
    DevGuard guard__(c);
    if (!c) return fail("plbm_step: null context");
    if (c->cfg.nranks > 1) return fail("plbm_step: this context is one slab of %d; drive it with plbm_step_local / plbm_halo_* / plbm_poisson_stage", c->cfg.nranks);
    auto eager = [&](int from, int to) -> int {
        for (int t = from; t < to; ++t) {
            if (one_step(c, want_fields && t == nsteps - 1, nullptr)) return 1;
            if (solve_poisson_in_loop(c, nullptr)) return 1;
        }
        return 0;
    };
    static const bool graphs_off = []() { const char* e = std::getenv("PLBM_NO_GRAPH"); return e && e[0] && e[0] != '0'; }();
    const long long cells = (long long)c->cfg.NX * c->geom.NYl;
    int done = 0;
    if (!graphs_off && !c->unfused && c->pop[0] && cells <= GRAPH_MAX_CELLS && nsteps >= GRAPH_MIN_STEPS) {
        if (eager(0, 2)) return 1;
        done = 2;
        const int pairs = (nsteps - 1 - done) / 2;              // the last step stays eager (it may have to write the moments)
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = eager(done, done + 2) ;                  // recorded, not run; host-side state advances as if it had run
        const cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
        if (rc || ce != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            return rc ? 1 : fail("cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
        }
        CUDA_TRY(cudaGraphInstantiate(&exec, graph, 0));
        for (int k = 0; k < pairs; ++k) CUDA_TRY(cudaGraphLaunch(exec, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));            // the executable graph must outlive its launches
        cudaGraphExecDestroy(exec);
        cudaGraphDestroy(graph);
        done += 2 * pairs;
        // host-side ping-pong state: the recorded pair advanced it once; the replays are whole pairs and leave it unchanged
        if (pairs == 0) return fail("plbm_step: internal: empty graph replay");
    }
    if (eager(done, nsteps)) return 1;
    if (want_fields && materialise_efield(c)) return 1;
    return 0;
}

int plbm_sync(plbm_ctx* c)
{
    DevGuard guard__(c);
    if (!c) return fail("plbm_sync: null context");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

int plbm_download_fields(plbm_ctx* c, double* const out[PLBM_NUM_FIELDS])
{
    DevGuard guard__(c);
    if (!c || !out) return fail("plbm_download_fields: null argument");
    if ((out[PLBM_F_EX] || out[PLBM_F_EY]) && materialise_efield(c)) return 1;
    const size_t bytes = sizeof(double) * (size_t)c->geom.NX * c->geom.NYl;
    for (int k = 0; k < PLBM_NUM_FIELDS; ++k) {
        if (!out[k]) continue;
        const double* src;
        if (k < 12) {
            if (!c->macro_valid) return fail("plbm_download_fields: moment fields were not requested from the last plbm_step");
            src = c->macro[k];
        } else if (k == PLBM_F_RHO_Q) src = c->rho_q;
        else if (k == PLBM_F_EX) src = c->Ex;
        else if (k == PLBM_F_EY) src = c->Ey;
        else src = c->phi;
        CUDA_TRY(cudaMemcpyAsync(out[k], src, bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

int plbm_fetch_begin(plbm_ctx* c, double* const out[PLBM_NUM_FIELDS])
{
    DevGuard guard__(c);
    if (!c || !out) return fail("plbm_fetch_begin: null argument");
    if (c->fetch_pending) return fail("plbm_fetch_begin: the previous fetch was not completed with plbm_fetch_wait");
    if ((out[PLBM_F_EX] || out[PLBM_F_EY]) && materialise_efield(c)) return 1;
    const size_t n = (size_t)c->geom.NX * c->geom.NYl, bytes = sizeof(double) * n;
    // first use: copy stream and events (plbm_frames_begin may have created them already), snapshot buffers, second macro set --
    // each guarded by its own pointer, so that a context that mixes plbm_frames_begin and plbm_fetch_begin has them all
    if (!c->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    if (!c->ev_snap) CUDA_TRY(cudaEventCreateWithFlags(&c->ev_snap, cudaEventDisableTiming));
    if (!c->ev_fetched) CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fetched, cudaEventDisableTiming));
    for (int k = 0; k < 4; ++k) if (!c->snap[k] && dev_alloc(c, &c->snap[k], n)) return 1;
    if (!c->unfused && !c->cfg.fields_only && !c->macro_sets[1][0]) {
        // the set the pending copy reads must not be the one the next step writes: from now on the steps alternate
        for (int k = 0; k < 12; ++k) if (dev_alloc(c, &c->macro_sets[1][k], n)) return 1;
    }
    bool any_macro = false;
    for (int k = 0; k < 12; ++k) any_macro |= (out[k] != nullptr);
    if (any_macro && !c->macro_valid) return fail("plbm_fetch_begin: moment fields were not requested from the last plbm_step");
    // rho_q / E / phi are overwritten by the next step: snapshot them on the compute stream (device copies)
    const double* live[4] = { c->rho_q, c->Ex, c->Ey, c->phi };
    for (int k = 0; k < 4; ++k)
        if (out[12 + k]) CUDA_TRY(cudaMemcpyAsync(c->snap[k], live[k], bytes, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaEventRecord(c->ev_snap, c->stream));
    CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_snap, 0));
    for (int k = 0; k < PLBM_NUM_FIELDS; ++k) {
        if (!out[k]) continue;
        const double* src = (k < 12) ? c->macro[k] : c->snap[k - 12];
        CUDA_TRY(cudaMemcpyAsync(out[k], src, bytes, cudaMemcpyDeviceToHost, c->copy_stream));
    }
    CUDA_TRY(cudaEventRecord(c->ev_fetched, c->copy_stream));
    if (c->unfused && any_macro)                           // single moment set on this path: the next sweep waits for the copy
        CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_fetched, 0));
    c->fetch_pending = true;
    return 0;
}

int plbm_frames_begin(plbm_ctx* c, float* const frames[PLBM_NUM_FRAMES], double* series)
{
    DevGuard guard__(c);
    if (!c) return fail("plbm_frames_begin: null context");
    if (c->fetch_pending) return fail("plbm_frames_begin: the previous fetch was not completed");
    if (!c->macro_valid) return fail("plbm_frames_begin: moment fields were not requested from the last plbm_step");
    const size_t n = (size_t)c->geom.NX * c->geom.NYl;
    if (!c->copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_snap, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fetched, cudaEventDisableTiming));
    }
    if (!c->frames) {
        if (dev_alloc(c, &c->frames, n * NUM_FRAMES)) return 1;
        if (dev_alloc(c, &c->series, (size_t)NUM_SERIES * NUM_POINTS)) return 1;
    }
    // narrowing runs on the compute stream right after the step, so the next step may overwrite every field
    if (frames) {
        const FrameFields f = { c->macro[PLBM_F_RHO_E], c->macro[PLBM_F_RHO_I], c->rho_q, c->macro[PLBM_F_UX_E], c->macro[PLBM_F_UY_E],
                                c->macro[PLBM_F_UX_I], c->macro[PLBM_F_UY_I], c->macro[PLBM_F_T_E], c->macro[PLBM_F_T_I], c->macro[PLBM_F_T_N] };
        CUDA_TRY(launch_frames(f, c->frames, n, c->stream));
    }
    if (series) {
        if (materialise_efield(c)) return 1;
        SeriesFields f;
        for (int s = 0; s < 3; ++s) {
            f.ux[s] = c->macro[2 * s]; f.uy[s] = c->macro[2 * s + 1]; f.T[s] = c->macro[6 + s]; f.rho[s] = c->macro[9 + s];
        }
        f.rho_q = c->rho_q; f.Ex = c->Ex; f.Ey = c->Ey;
        CUDA_TRY(launch_series(f, c->series, c->cfg.NX, c->cfg.NY, c->cfg.y0, c->geom.NYl, c->stream));
    }
    CUDA_TRY(cudaEventRecord(c->ev_snap, c->stream));
    CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_snap, 0));
    if (frames)
        for (int k = 0; k < NUM_FRAMES; ++k)
            if (frames[k]) CUDA_TRY(cudaMemcpyAsync(frames[k], c->frames + (size_t)k * n, sizeof(float) * n, cudaMemcpyDeviceToHost, c->copy_stream));
    if (series) CUDA_TRY(cudaMemcpyAsync(series, c->series, sizeof(double) * NUM_SERIES * NUM_POINTS, cudaMemcpyDeviceToHost, c->copy_stream));
    CUDA_TRY(cudaEventRecord(c->ev_fetched, c->copy_stream));
    c->fetch_pending = true;                                   // one frame buffer: the next narrowing comes after plbm_fetch_wait
    return 0;
}

int plbm_pin_host(void* p, size_t bytes)
{
    if (!p || !bytes) return fail("plbm_pin_host: empty range");
    CUDA_TRY(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
    return 0;
}
int plbm_unpin_host(void* p)
{
    if (!p) return 0;
    CUDA_TRY(cudaHostUnregister(p));
    return 0;
}

int plbm_fetch_wait(plbm_ctx* c)
{
    DevGuard guard__(c);
    if (!c) return fail("plbm_fetch_wait: null context");
    if (!c->fetch_pending) return 0;
    CUDA_TRY(cudaEventSynchronize(c->ev_fetched));
    c->fetch_pending = false;
    return 0;
}

int plbm_step_timed(plbm_ctx* c, int nsteps, int want_fields, float* ms_total, float* ms_k1, float* ms_poisson, long long* launches)
{
    DevGuard guard__(c);
    if (!c) return fail("plbm_step_timed: null context");
    if (nsteps < 1 || nsteps > 100000) return fail("plbm_step_timed: nsteps out of range");
    if (c->cfg.nranks > 1) return fail("plbm_step_timed: single-slab contexts only");
    const size_t need = (size_t)2 * nsteps + 1;
    while (c->events.size() < need) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreate(&e));
        c->events.push_back(e);
    }
    long long nl = 0;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaEventRecord(c->events[0], c->stream));
    for (int t = 0; t < nsteps; ++t) {
        if (one_step(c, want_fields && t == nsteps - 1, &nl)) return 1;
        CUDA_TRY(cudaEventRecord(c->events[2 * t + 1], c->stream));
        if (solve_poisson_in_loop(c, &nl)) return 1;
        if (want_fields && t == nsteps - 1 && materialise_efield(c, &nl)) return 1;
        CUDA_TRY(cudaEventRecord(c->events[2 * t + 2], c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    float total = 0.f, k1 = 0.f, ps = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&total, c->events[0], c->events[2 * nsteps]));
    for (int t = 0; t < nsteps; ++t) {
        float a = 0.f, b = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&a, c->events[2 * t], c->events[2 * t + 1]));
        CUDA_TRY(cudaEventElapsedTime(&b, c->events[2 * t + 1], c->events[2 * t + 2]));
        k1 += a; ps += b;
    }
    if (ms_total) *ms_total = total;
    if (ms_k1) *ms_k1 = k1;
    if (ms_poisson) *ms_poisson = ps;
    if (launches) *launches = nl;
    return 0;
}

int plbm_host_solve_poisson(plbm_ctx* c, const double* rho_q, double* Ex, double* Ey)
{
    DevGuard guard__(c);
    if (!c || !rho_q || !Ex || !Ey) return fail("plbm_host_solve_poisson: null argument");
    const size_t bytes = sizeof(double) * (size_t)c->geom.NX * c->geom.NYl;
    CUDA_TRY(cudaMemcpyAsync(c->rho_q, rho_q, bytes, cudaMemcpyHostToDevice, c->stream));
    c->e_stale = false;
    CUDA_TRY(cudaMemcpyAsync(c->Ex, Ex, bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->Ey, Ey, bytes, cudaMemcpyHostToDevice, c->stream));
    if (solve_poisson(c, nullptr)) return 1;
    CUDA_TRY(cudaMemcpyAsync(Ex, c->Ex, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(Ey, c->Ey, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

int plbm_host_poisson_config(plbm_ctx* c, int poisson_type, int bc_type, double omega_sor)
{
    DevGuard guard__(c);
    if (!c) return fail("plbm_host_poisson_config: null context");
    if (c->pop[0] || c->unfused) return fail("plbm_host_poisson_config: only for fields-only contexts (the time loop keeps the type it was created with)");
    if (poisson_type < PLBM_POISSON_NONE || poisson_type > PLBM_POISSON_NPS) return fail("plbm_host_poisson_config: Poisson type %d", poisson_type);
    if (bc_type != PLBM_BC_PERIODIC && bc_type != PLBM_BC_BOUNCEBACK) return fail("plbm_host_poisson_config: boundary type %d", bc_type);
    c->cfg.poisson_type = poisson_type;
    c->cfg.bc_type = bc_type;
    c->cfg.omega_sor = omega_sor;
    return 0;
}

int plbm_host_poisson_solver(plbm_ctx* c, int poisson_type, const double* rho_q)
{
    DevGuard guard__(c);
    if (!c || !rho_q) return fail("plbm_host_poisson_solver: null argument");
    const size_t bytes = sizeof(double) * (size_t)c->geom.NX * c->geom.NYl;
    if (poisson_first_call(c)) return 1;
    CUDA_TRY(cudaMemcpyAsync(c->rho_q, rho_q, bytes, cudaMemcpyHostToDevice, c->stream));
    if (poisson_solver(c, poisson_type, nullptr)) return 1;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

int plbm_host_efield(plbm_ctx* c, int bc_type, double* Ex, double* Ey)
{
    DevGuard guard__(c);
    if (!c || !Ex || !Ey) return fail("plbm_host_efield: null argument");
    const size_t bytes = sizeof(double) * (size_t)c->geom.NX * c->geom.NYl;
    if (poisson_first_call(c)) return 1;
    c->e_stale = false;
    CUDA_TRY(cudaMemcpyAsync(c->Ex, Ex, bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->Ey, Ey, bytes, cudaMemcpyHostToDevice, c->stream));
    if (poisson_efield(c, bc_type, nullptr)) return 1;
    CUDA_TRY(cudaMemcpyAsync(Ex, c->Ex, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(Ey, c->Ey, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// The cut.  K1's time per cell depends on the plasma the cell holds (k1_kernel.cuh: cells whose electrons and ions are empty run
// at the memory system's pace, dense cells on the FP64 pipe), and LBmethod's lattices start from the reference's initial condition,
// electrons and ions in the rows NY/4 < y < 3NY/4 only (src/plasma.cpp:128-143).  An equal split of the rows then makes the slabs
// without plasma wait for the slabs with it at the first barrier of every step (424 of 3548 us at 8192^2 on 4 B200s, 227 of 1910 on 8).
// The rows of the central block therefore weigh 1 + alpha: cut r is where the weight reaches r/R of the total, rounded to an even
// row.  alpha = 0.17 from the measured costs per cell at 8192^2 (133.7 ps empty, 184 ps dense, half of a block row dense, 16.7 ps of
// row passes of the spectral solve); PLBM_SLAB_ALPHA overrides, 0 gives the equal split.  Results do not depend on the cut.
static double slab_alpha()
{
    static const double a = []() {
        const char* e = std::getenv("PLBM_SLAB_ALPHA");
        double v = e ? std::atof(e) : 0.17;
        return (v >= 0.0 && v <= 4.0) ? v : 0.0;
    }();
    return a;
}
static int slab_cut(int NY, int r, int R)
{
    if (r <= 0) return 0;
    if (r >= R) return NY;
    const double alpha = slab_alpha();
    const long long pairs = NY / 2;
    if (alpha == 0.0) return (int)(2 * ((pairs * r) / R));
    const double b0 = NY / 4 + 1, b1 = (3 * NY) / 4, wb = (1.0 + alpha) * (b1 - b0);
    const double t = (NY + alpha * (b1 - b0)) * r / R;
    const double y = (t <= b0) ? t : (t <= b0 + wb ? b0 + (t - b0) / (1.0 + alpha) : t - alpha * (b1 - b0));
    long long c = 2 * (long long)(y / 2.0 + 0.5);
    if (c < 0) c = 0;
    if (c > NY) c = NY;
    return (int)c;
}

int plbm_slab_of(int NY, int rank, int nranks, int* y0, int* ny_local)
{
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail("plbm_slab_of: rank %d of %d", rank, nranks);
    if (nranks == 1) { if (y0) *y0 = 0; if (ny_local) *ny_local = NY; return 0; }
    if (NY % 2) return fail("plbm_slab_of: NY must be even to cut slabs (the spectral solve transforms rows in pairs)");
    const long long pairs = NY / 2;
    // the weighted cut must leave every slab at least two rows; otherwise (small lattices) the equal split
    bool weighted = slab_alpha() > 0.0;
    for (int r = 0; weighted && r < nranks; ++r) if (slab_cut(NY, r + 1, nranks) - slab_cut(NY, r, nranks) < 2) weighted = false;
    const int a = weighted ? slab_cut(NY, rank, nranks) : (int)(2 * ((pairs * rank) / nranks));
    const int b = weighted ? slab_cut(NY, rank + 1, nranks) : (int)(2 * ((pairs * (rank + 1)) / nranks));
    if (b - a < 2) return fail("plbm_slab_of: %d rows cannot feed %d slabs", NY, nranks);
    if (y0) *y0 = a;
    if (ny_local) *ny_local = b - a;
    return 0;
}

int plbm_step_local(plbm_ctx* c, int want_fields)
{
    DevGuard guard__(c);
    if (!c) return fail("plbm_step_local: null context");
    return one_step(c, want_fields != 0, nullptr);
}

int plbm_halo_pack(plbm_ctx* c)
{
    DevGuard guard__(c);
    if (!c || c->cfg.nranks < 2) return fail("plbm_halo_pack: needs a multi-slab context");
    if (c->halo_fresh) return 0;            // initialise / upload filled the halo rows: nothing to exchange (see plbm_upload_state)
    CUDA_TRY(launch_halo_pack(c->pop[c->cur], c->halo_send_lo, c->halo_send_hi, c->geom, c->stream));
    return 0;
}

int plbm_halo_unpack(plbm_ctx* c)
{
    DevGuard guard__(c);
    if (!c || c->cfg.nranks < 2) return fail("plbm_halo_unpack: needs a multi-slab context");
    if (c->halo_fresh) return 0;            // initialise / upload filled the halo rows: nothing to exchange (see plbm_upload_state)
    const size_t hoff = (size_t)c->halo_par * 18 * c->cfg.NX;
    CUDA_TRY(launch_halo_unpack(c->pop[c->cur], c->halo_recv_lo + hoff, c->halo_recv_hi + hoff, c->geom, c->stream));
    return 0;
}

// P2 through peer memory.  From four slabs on, copier CTAs gather the columns beside the transforms (poisson_cols.cu): on 8 B200s
// 161 vs 200 us at 6144^2 and 235 vs 296 us at 8192^2.  Not with two slabs (half of every column is local and the kernel whose CTAs
// load their own columns over NVLink is as fast or faster: 131 vs 144 us at 3072^2, 556 vs 537 us at 8192^2 on 2 B200s) and not
// beyond 8192 (768-thread transforms: 989 vs 887 us at 12288^2 on 8 B200s), nor up to 4096.  PLBM_P2_GATHER=1 / 0 forces one or the other.
static cudaError_t launch_p2_peer(plbm_ctx* c, cudaStream_t stream)
{
    static const int forced = []() { const char* e = std::getenv("PLBM_P2_GATHER"); return (e && e[0]) ? (e[0] != '0' ? 1 : 0) : -1; }();
    // (copiers pay where a transform CTA fills an SM: lengths above 4096; at 4096 two CTAs share an SM and cover each other's loads:
    // 114 vs 158 us at 4096^2 on 4 B200s, against 386 vs 423 us at 8192^2)
    const bool gather = forced >= 0 ? forced == 1 : (c->cfg.nranks >= 4 && c->fft.col.threads > 256 && c->fft.col.threads <= 512);
    if (gather && c->fft.p2_flags) return launch_poisson_cols_gather(c->fft, stream, c->peer_t1);
    return launch_poisson_cols(c->fft, stream, &c->peer_t1);
}

int plbm_poisson_stage(plbm_ctx* c, int stage)
{
    DevGuard guard__(c);
    if (!c) return fail("plbm_poisson_stage: null context");
    if (poisson_first_call(c)) return 1;
    const int type = c->cfg.poisson_type;
    if (type == PLBM_POISSON_NONE) return 0;
    if (type != PLBM_POISSON_FFT || c->cfg.bc_type != PLBM_BC_PERIODIC) return fail("plbm_poisson_stage: only the periodic spectral solve runs on several slabs");
    switch (stage) {
    case 0: CUDA_TRY(launch_poisson_rows_fwd(c->fft, c->rho_q, c->stream)); return 0;
    case 1: CUDA_TRY(launch_poisson_cols(c->fft, c->stream)); return 0;
    case 2: CUDA_TRY(launch_poisson_rows_inv(c->fft, c->phi, c->stream)); return 0;
    case 5:                                  // stage 2 + the boundary rows of phi written into the neighbours as well (peer memory)
        if (!c->peers) return fail("plbm_poisson_stage(5): peer memory is not attached (plbm_peer_attach)");
        CUDA_TRY(launch_poisson_rows_inv(c->fft, c->phi, c->stream, c->down_phi_above, c->up_phi_below));
        return 0;
    case 3:                                  // field reconstruction: implicit in phi for K1, materialised on demand
        if (c->unfused || !c->pop[0]) return poisson_efield(c, PLBM_BC_PERIODIC, nullptr);
        c->e_stale = true;
        return 0;
    case 4:
        if (!c->peers) return fail("plbm_poisson_stage(4): peer memory is not attached (plbm_peer_attach)");
        CUDA_TRY(launch_p2_peer(c, c->stream));
        return 0;
    default: return fail("plbm_poisson_stage: stage %d", stage);
    }
}

namespace {
// Second stream, events and buffers of the pipelined peer step, on first use.
int ensure_pipeline(plbm_ctx* c)
{
    if (c->pstream) return 0;
    const size_t n = (size_t)c->cfg.NX * c->geom.NYl;
    int lo = 0, hi = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_TRY(cudaStreamCreateWithPriority(&c->pstream, cudaStreamNonBlocking, hi));   // its short kernels go first when a slot frees up
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_pop, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_phi, cudaEventDisableTiming));
    if (dev_alloc(c, &c->rho_q_next, n)) return 1;
    if (dev_alloc(c, &c->phi_buf[1], n)) return 1;
    CUDA_TRY(cudaMemsetAsync(c->phi_buf[1], 0, sizeof(double) * n, c->stream));
    return 0;
}

int ensure_peer_flags(plbm_ctx* c)
{
    if (c->flags) return 0;
    if (dev_alloc(c, &c->flags, 2 * PLBM_MAX_RANKS)) return 1;          // [0..MAX): barriers of the main stream, [MAX..2MAX): of the Poisson stream
    if (dev_alloc(c, &c->peer_timeout, 1)) return 1;
    CUDA_TRY(cudaMemset(c->flags, 0, sizeof(unsigned long long) * 2 * PLBM_MAX_RANKS));
    CUDA_TRY(cudaMemset(c->peer_timeout, 0, sizeof(int)));
    return 0;
}
} // namespace

int plbm_peer_export(plbm_ctx* c, void* blob)
{
    DevGuard guard__(c);
    if (!c || !blob) return fail("plbm_peer_export: null argument");
    if (c->cfg.nranks < 2 || !c->fft.T1) return fail("plbm_peer_export: needs a multi-slab context with the spectral Poisson solve");
    if (ensure_peer_flags(c)) return 1;
    PeerBlob b;
    std::memset(&b, 0, sizeof(b));
    const void* bufs[PEER_NBUF] = { c->fft.T1, c->flags, c->halo_recv_lo, c->halo_recv_hi, c->phi_below_base, c->phi_above_base };
    for (int i = 0; i < PEER_NBUF; ++i) {
        CUDA_TRY(cudaIpcGetMemHandle(&b.handle[i], const_cast<void*>(bufs[i])));
        if (allocation_offset(bufs[i], &b.offset[i])) return 1;
    }
    b.rank = c->cfg.rank; b.nyl = c->geom.NYl;
    std::memset(blob, 0, PLBM_PEER_BLOB_BYTES);
    std::memcpy(blob, &b, sizeof(b));
    return 0;
}

int plbm_peer_attach(plbm_ctx* c, const void* blobs)
{
    DevGuard guard__(c);
    if (!c || !blobs) return fail("plbm_peer_attach: null argument");
    if (!c->flags) return fail("plbm_peer_attach: call plbm_peer_export first");
    if (c->peers) return 0;
    const int R = c->cfg.nranks, me = c->cfg.rank, up = (me + 1) % R, down = (me + R - 1) % R;
    for (int s = 0; s < R; ++s) {
        PeerBlob b;
        std::memcpy(&b, (const char*)blobs + (size_t)s * PLBM_PEER_BLOB_BYTES, sizeof(b));
        if (b.rank != s || b.nyl != c->slab_y0[s + 1] - c->slab_y0[s]) return fail("plbm_peer_attach: blob %d does not describe slab %d", s, s);
        void* ptr[PEER_NBUF] = {};
        if (s == me) {
            void* mine[PEER_NBUF] = { c->fft.T1, c->flags, c->halo_recv_lo, c->halo_recv_hi, c->phi_below_base, c->phi_above_base };
            for (int i = 0; i < PEER_NBUF; ++i) ptr[i] = mine[i];
        } else {
            for (int i = 0; i < PEER_NBUF; ++i) {
                // buffers may share an underlying allocation (cudaMalloc sub-allocates): one mapping per distinct handle
                void* base = nullptr;
                for (int j = 0; j < i && !base; ++j)
                    if (std::memcmp(&b.handle[i], &b.handle[j], sizeof(cudaIpcMemHandle_t)) == 0) base = (char*)ptr[j] - b.offset[j];
                if (!base) {
                    CUDA_TRY(cudaIpcOpenMemHandle(&base, b.handle[i], cudaIpcMemLazyEnablePeerAccess));
                    c->peer_mapped[PEER_NBUF * s + i] = base;
                }
                ptr[i] = (char*)base + b.offset[i];
            }
        }
        c->peer_t1.t1[s] = (cpx*)ptr[0];
        c->peer_flags[s] = (unsigned long long*)ptr[1];
        if (s == up) { c->up_halo_recv_lo = (double*)ptr[2]; c->up_phi_below_base = (double*)ptr[4]; }
        if (s == down) { c->down_halo_recv_hi = (double*)ptr[3]; c->down_phi_above_base = (double*)ptr[5]; }
    }
    set_phi_parity(c, c->phi_par);
    c->peers = true;
    return 0;
}

// Same wiring for slabs that live in ONE process (plbm_group): no IPC, the sibling contexts' pointers are used directly after
// peer access between the devices has been enabled.  all[s] is the context of slab s (all[rank] == c).
int plbm_peer_attach_local(plbm_ctx* c, plbm_ctx* const* all)
{
    DevGuard guard__(c);
    if (!c || !all) return fail("plbm_peer_attach_local: null argument");
    if (c->cfg.nranks < 2 || !c->fft.T1) return fail("plbm_peer_attach_local: needs a multi-slab context with the spectral Poisson solve");
    if (c->peers) return 0;
    const int R = c->cfg.nranks, me = c->cfg.rank, up = (me + 1) % R, down = (me + R - 1) % R;
    for (int s = 0; s < R; ++s) {
        plbm_ctx* o = all[s];
        if (!o || o->cfg.rank != s || o->cfg.nranks != R || !o->fft.T1) return fail("plbm_peer_attach_local: context %d is not slab %d of %d", s, s, R);
        if (!o->flags) return fail("plbm_peer_attach_local: slab %d has no flag buffer yet (plbm_peer_prepare_local on every slab first)", s);
        if (s != me && o->device != c->device) {
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, c->device, o->device));
            if (!can) return fail("plbm_peer_attach_local: device %d cannot access device %d", c->device, o->device);
            const cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return fail("cudaDeviceEnablePeerAccess(%d) failed: %s", o->device, cudaGetErrorString(e));
        }
        c->peer_t1.t1[s] = o->fft.T1;
        c->peer_flags[s] = o->flags;
        if (s == up) { c->up_halo_recv_lo = o->halo_recv_lo; c->up_phi_below_base = o->phi_below_base; }
        if (s == down) { c->down_halo_recv_hi = o->halo_recv_hi; c->down_phi_above_base = o->phi_above_base; }
    }
    set_phi_parity(c, c->phi_par);
    c->peers = true;
    return 0;
}
int plbm_peer_prepare_local(plbm_ctx* c)
{
    DevGuard guard__(c);
    if (!c) return fail("plbm_peer_prepare_local: null context");
    return ensure_peer_flags(c);
}

namespace {

int peer_barrier_on(plbm_ctx* c, cudaStream_t stream, int set)
{
    PeerFlags pf;
    for (int s = 0; s < PLBM_MAX_RANKS; ++s) pf.f[s] = c->peer_flags[s] ? c->peer_flags[s] + set * PLBM_MAX_RANKS : nullptr;
    const unsigned long long epoch = set ? ++c->epoch_b : ++c->epoch;
    peer_barrier_kernel<<<1, 32, 0, stream>>>(pf, c->cfg.rank, c->cfg.nranks, epoch, c->peer_timeout);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Stage times of plbm_step_peer (PLBM_PEER_STAGES slots): CUDA events on the stream the stage runs on.
struct StageClock {
    static constexpr int TIMED = 32;
    float* out;
    int nsteps, first_timed;
    std::vector<cudaEvent_t> ev;              // [timed step][slot][begin, end]
    StageClock(float* o, int n) : out(o), nsteps(n), first_timed(n > TIMED ? n - TIMED : 0)
    {
        if (out) ev.assign((size_t)(nsteps - first_timed) * PLBM_PEER_STAGES * 2, nullptr);
    }
    ~StageClock() { for (auto e : ev) if (e) cudaEventDestroy(e); }
    int mark(int t, int slot, int end, cudaStream_t s)
    {
        if (!out || t < first_timed) return 0;
        cudaEvent_t& e = ev[((size_t)(t - first_timed) * PLBM_PEER_STAGES + slot) * 2 + end];
        if (!e) CUDA_TRY(cudaEventCreate(&e));
        CUDA_TRY(cudaEventRecord(e, s));
        return 0;
    }
    void collect()
    {
        if (!out) return;
        for (size_t k = 0; k + 1 < ev.size(); k += 2) {
            float ms = 0.f;
            if (ev[k] && ev[k + 1] && cudaEventElapsedTime(&ms, ev[k], ev[k + 1]) == cudaSuccess) out[(k / 2) % PLBM_PEER_STAGES] += ms;
        }
    }
};
enum { ST_K1 = 0, ST_PUSH, ST_HALO_BAR, ST_UNPACK, ST_PHI_WAIT, ST_PULL, ST_P1, ST_BAR1, ST_P2, ST_BAR2, ST_P3, ST_BAR3 };

#define STAGE(slot, stream, call) (clk.mark(t, slot, 0, stream) || (call) || clk.mark(t, slot, 1, stream))

// K1, halo push, P1, BARRIER, halo unpack, P2 over peer memory, BARRIER, P3 + phi rows into the neighbours, BARRIER -- one stream
// (the sequence documented in plbm.h; the default)
int step_peer_sequential(plbm_ctx* c, int nsteps, bool want_fields, StageClock& clk)
{
    cudaStream_t A = c->stream;
    int rc = 0;
    for (int t = 0; t < nsteps && !rc; ++t) {
        rc = STAGE(ST_K1, A, one_step(c, want_fields && t == nsteps - 1, nullptr))
          || STAGE(ST_PUSH, A, plbm_halo_push(c))
          || STAGE(ST_P1, A, plbm_poisson_stage(c, 0))
          || STAGE(ST_BAR1, A, peer_barrier_on(c, A, 0))
          || STAGE(ST_UNPACK, A, plbm_halo_unpack(c))
          || STAGE(ST_P2, A, plbm_poisson_stage(c, 4))
          || STAGE(ST_BAR2, A, peer_barrier_on(c, A, 0))
          || STAGE(ST_P3, A, plbm_poisson_stage(c, 5))
          || STAGE(ST_BAR3, A, peer_barrier_on(c, A, 0))
          || plbm_poisson_stage(c, 3);
    }
    return rc;
}

// The same step with the Poisson solve OFF the critical path.  rho_q of step t depends only on the populations step t-1 wrote
// (charge_pull.cu), so phi of step t -- which K1 of step t+1 needs -- can be computed while K1 of step t runs:
//   stream A:  [pop of t-1 complete] -> K1 t (reads phi t-1) -> halo push -> BARRIER_A -> halo unpack -> wait for phi t -> ...
//   stream B:  [pop of t-1 complete] -> charge pull -> P1 -> BARRIER_B -> P2 over peer memory -> BARRIER_B -> P3 into the OTHER
//              potential + boundary rows into the neighbours' other copies -> BARRIER_B -> [phi t ready]
// Hazards.  K1 t reads potential `par` while P3 t writes `par^1`; the flip happens on every slab after both streams have met.  A
// neighbour's P3 t+2 overwrites my copies `par` again only after BARRIER_B (2nd of t+2), which I join after my K1 t+1 has
// finished reading them (stream B of t+2 starts from the event recorded after K1 t+1 and its halo exchange).  The charge pull of t
// reads the planes K1 t+1 will overwrite; K1 t+1 starts after "phi t ready".  Halo receive buffers alternate (halo_par): the push of
// t+1 goes into the set the neighbour is not unpacking, the push of t+2 follows BARRIER_A of t+1, which the neighbour joins
// after its unpack of t.  T1 is reused inside stream B in the order the sequential version already relies on.
int step_peer_pipelined(plbm_ctx* c, int nsteps, bool want_fields, StageClock& clk)
{
    if (ensure_pipeline(c)) return 1;
    if (poisson_first_call(c)) return 1;
    cudaStream_t A = c->stream, B = c->pstream;
    const LbmGeom& g = c->geom;
    int rc = 0;
    for (int t = 0; t < nsteps && !rc; ++t) {
        const int alt = c->phi_par ^ 1;
        const size_t NX = (size_t)c->cfg.NX;
        // A -> B: the populations of the previous step are complete, halo rows included
        CUDA_TRY(cudaEventRecord(c->ev_pop, A));
        CUDA_TRY(cudaStreamWaitEvent(B, c->ev_pop, 0));
        rc = STAGE(ST_PULL, B, (launch_charge_pull(c->pop[c->cur], c->rho_q_next, g, c->consts.q, c->mass, B) != cudaSuccess ? fail("charge pull launch failed") : 0))
          || STAGE(ST_P1, B, (launch_poisson_rows_fwd(c->fft, c->rho_q_next, B) != cudaSuccess ? fail("P1 launch failed") : 0))
          || STAGE(ST_BAR1, B, peer_barrier_on(c, B, 1))
          || STAGE(ST_P2, B, (launch_p2_peer(c, B) != cudaSuccess ? fail("P2 launch failed") : 0))
          || STAGE(ST_BAR2, B, peer_barrier_on(c, B, 1))
          || STAGE(ST_P3, B, (launch_poisson_rows_inv(c->fft, c->phi_buf[alt], B, c->down_phi_above_base + alt * NX, c->up_phi_below_base + alt * NX) != cudaSuccess
                              ? fail("P3 launch failed") : 0))
          || STAGE(ST_BAR3, B, peer_barrier_on(c, B, 1));
        if (rc) break;
        CUDA_TRY(cudaEventRecord(c->ev_phi, B));
        // A: the step itself, beside all of the above
        rc = STAGE(ST_K1, A, one_step(c, want_fields && t == nsteps - 1, nullptr))
          || STAGE(ST_PUSH, A, plbm_halo_push(c))
          || STAGE(ST_HALO_BAR, A, peer_barrier_on(c, A, 0))
          || STAGE(ST_UNPACK, A, plbm_halo_unpack(c));
        if (rc) break;
        c->halo_par ^= 1;
        // B -> A: the next K1 (and any download) sees phi of this step; the wait is what is left of the solve on the critical path
        if (clk.mark(t, ST_PHI_WAIT, 0, A)) return 1;
        CUDA_TRY(cudaStreamWaitEvent(A, c->ev_phi, 0));
        if (clk.mark(t, ST_PHI_WAIT, 1, A)) return 1;
        set_phi_parity(c, alt);
        c->e_stale = true;
    }
    return rc;
}
#undef STAGE

} // namespace

// One whole time step of a slab with peers attached, `nsteps` times, in one call (no host round trip per kernel).  Default: the
// sequential sequence; PLBM_PEER_PIPELINE=1 in the environment selects the pipelined one (bit-identical; on 2 B200s at 3072^2 it is
// slower, 1.44 vs 1.32 ms per step: the Poisson kernels take the SMs they run on away from K1, which gets 300 us slower).  stage_ms, if given, holds
// PLBM_PEER_STAGES floats and receives the accumulated device time of each stage over at most the last 32 steps.
int plbm_step_peer(plbm_ctx* c, int nsteps, int want_fields, float* stage_ms)
{
    DevGuard guard__(c);
    if (!c || !c->peers) return fail("plbm_step_peer: peer memory is not attached");
    if (nsteps < 0) return fail("plbm_step_peer: nsteps = %d", nsteps);
    const char* e = std::getenv("PLBM_PEER_PIPELINE");
    const bool pipelined = (e && e[0] == '1');       // measured slower on 2 B200s (profiles/r2_summary.md): opt-in
    StageClock clk(stage_ms, nsteps);
    int rc = pipelined ? step_peer_pipelined(c, nsteps, want_fields != 0, clk) : step_peer_sequential(c, nsteps, want_fields != 0, clk);
    if (!rc && stage_ms) {
        if (cudaStreamSynchronize(c->stream) != cudaSuccess || (c->pstream && cudaStreamSynchronize(c->pstream) != cudaSuccess))
            rc = fail("plbm_step_peer: stream synchronisation failed");
        else clk.collect();
    }
    return rc;
}

// halo rows straight into the neighbours' receive buffers / boundary rows of phi into the neighbours' copies
int plbm_halo_push(plbm_ctx* c)
{
    DevGuard guard__(c);
    if (!c || !c->peers) return fail("plbm_halo_push: peer memory is not attached");
    if (c->halo_fresh) return 0;            // initialise / upload filled the halo rows: nothing to exchange (see plbm_upload_state)
    const size_t hoff = (size_t)c->halo_par * 18 * c->cfg.NX;
    CUDA_TRY(launch_halo_pack(c->pop[c->cur], c->down_halo_recv_hi + hoff, c->up_halo_recv_lo + hoff, c->geom, c->stream));
    return 0;
}
int plbm_phi_rows_push(plbm_ctx* c)
{
    DevGuard guard__(c);
    if (!c || !c->peers) return fail("plbm_phi_rows_push: peer memory is not attached");
    const size_t row = sizeof(double) * c->cfg.NX;
    CUDA_TRY(cudaMemcpyAsync(c->up_phi_below, c->phi + (size_t)(c->geom.NYl - 1) * c->cfg.NX, row, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->down_phi_above, c->phi, row, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

int plbm_peer_detach(plbm_ctx* c)
{
    DevGuard guard__(c);
    if (!c) return fail("plbm_peer_detach: null context");
    if (c->stream) CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (void*& m : c->peer_mapped) {
        if (m) CUDA_TRY(cudaIpcCloseMemHandle(m));
        m = nullptr;
    }
    c->peers = false;
    return 0;
}

int plbm_peer_barrier(plbm_ctx* c)
{
    DevGuard guard__(c);
    if (!c || !c->peers) return fail("plbm_peer_barrier: peer memory is not attached");
    return peer_barrier_on(c, c->stream, 0);
}

int plbm_peer_check(plbm_ctx* c)
{
    DevGuard guard__(c);
    if (!c || !c->peers) return 0;
    int t = 0;
    CUDA_TRY(cudaMemcpyAsync(&t, c->peer_timeout, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (t) return fail("a peer-memory barrier timed out: another slab's process stopped making progress");
    return 0;
}

int plbm_exchange_info(plbm_ctx* c, plbm_exchange* o)
{
    DevGuard guard__(c);
    if (!c || !o) return fail("plbm_exchange_info: null argument");
    std::memset(o, 0, sizeof(*o));
    const int R = c->cfg.nranks, nh = c->cfg.NY / 2 + 1;
    o->nranks = R; o->rank = c->cfg.rank;
    o->halo_send_lo = c->halo_send_lo; o->halo_send_hi = c->halo_send_hi;
    o->halo_recv_lo = c->halo_recv_lo; o->halo_recv_hi = c->halo_recv_hi;
    o->halo_count = (long long)18 * c->cfg.NX;
    o->phi_first_row = c->phi; o->phi_last_row = c->phi + (size_t)(c->geom.NYl - 1) * c->cfg.NX;
    o->phi_below = c->phi_below; o->phi_above = c->phi_above;
    o->phi_count = c->cfg.NX;
    o->t1 = c->fft.T1; o->t2 = c->fft.T2;
    for (int r = 0; r <= R; ++r) { o->slab_y0[r] = (R > 1) ? c->slab_y0[r] : (r ? c->cfg.NY : 0); o->slab_k0[r] = c->fft.T1 ? c->slab_k0[r] : 0; }
    (void)nh;
    return 0;
}

int plbm_local_rows(const plbm_ctx* c, int* y0, int* ny_local)
{
    DevGuard guard__(c);
    if (!c) return fail("plbm_local_rows: null context");
    if (y0) *y0 = c->cfg.y0;
    if (ny_local) *ny_local = c->geom.NYl;
    return 0;
}

long long plbm_device_bytes(const plbm_ctx* c) { return c ? c->bytes : 0; }
void* plbm_stream(plbm_ctx* c) { return c ? (void*)c->stream : nullptr; }

} // extern "C"
