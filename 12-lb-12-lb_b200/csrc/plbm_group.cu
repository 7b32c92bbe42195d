// plbm_group.cu -- several GPUs of one node behind ONE host object (include/plbm.h: plbm_group_*).
//
// The reference is a single C++ program (src/main_plasma.cpp:59-73 constructs one LBmethod and calls Run_simulation()); a
// group lets that program use every GPU of the box without becoming several processes: one context per device, the lattice
// cut into y-slabs by the library's own rule, the slabs wired through peer memory (plbm_peer_attach_local), and one host
// thread per device that issues the slab's step sequence (plbm_step_peer).  Between the slabs nothing crosses the host: halo
// rows, the spectral transposes and the boundary rows of phi are written / read in place over NVLink by the kernels, ordered
// by the flag barriers of plbm_peer_barrier.
#include "../../include/plbm.h"

#include <cuda_runtime.h>

#include "poisson_fft.h"
using plbm::PLBM_MAX_RANKS;

int plbm_set_error(const char* msg);             // plbm_api.cu: sets the calling thread's plbm_last_error() text, returns 1

#include <string>
#include <thread>
#include <vector>

struct plbm_group {
    std::vector<plbm_ctx*> member;
    std::vector<int> y0, nyl;
    int NX = 0, NY = 0;
};

namespace {

int group_fail(const std::string& msg) { return plbm_set_error(msg.c_str()); }

// Runs fn(slab) for every slab, each on its own host thread (the launches of one slab never wait for another slab's host
// code; the device-side barriers need every slab's kernels enqueued).  Returns the first failing slab's error.
template <class F>
int for_each_slab(plbm_group* g, F fn)
{
    const int n = (int)g->member.size();
    std::vector<int> rc(n, 0);
    std::vector<std::string> err(n);
    auto body = [&](int s) {
        rc[s] = fn(s);
        if (rc[s]) err[s] = plbm_last_error();          // the error text is thread-local
    };
    std::vector<std::thread> th;
    th.reserve(n > 0 ? n - 1 : 0);
    for (int s = 1; s < n; ++s) th.emplace_back(body, s);
    body(0);
    for (auto& t : th) t.join();
    for (int s = 0; s < n; ++s)
        if (rc[s]) return group_fail("slab " + std::to_string(s) + ": " + err[s]);
    return 0;
}

} // namespace

extern "C" {

int plbm_group_create(const plbm_config* cfg, int ndevices, plbm_group** out)
{
    if (!cfg || !out) return group_fail("plbm_group_create: null argument");
    *out = nullptr;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) return group_fail("plbm_group_create: no CUDA device (this library has no CPU path)");
    const int first = cfg->device >= 0 ? cfg->device : 0;
    if (ndevices < 2 || ndevices > PLBM_MAX_RANKS) return group_fail("plbm_group_create: ndevices = " + std::to_string(ndevices) + " (2.." + std::to_string(PLBM_MAX_RANKS) + ")");
    if (first + ndevices > have) return group_fail("plbm_group_create: devices " + std::to_string(first) + ".." + std::to_string(first + ndevices - 1) + " requested, " + std::to_string(have) + " present");
    if (cfg->poisson_type != PLBM_POISSON_FFT || cfg->bc_type != PLBM_BC_PERIODIC)
        return group_fail("plbm_group_create: several GPUs run the periodic spectral configuration only");
    plbm_group* g = new plbm_group();
    g->NX = cfg->NX; g->NY = cfg->NY;
    g->member.assign(ndevices, nullptr);
    g->y0.assign(ndevices, 0); g->nyl.assign(ndevices, 0);
    int rc = 0;
    for (int s = 0; s < ndevices && !rc; ++s) {
        plbm_config c = *cfg;
        c.rank = s; c.nranks = ndevices; c.device = first + s;
        rc = plbm_slab_of(cfg->NY, s, ndevices, &g->y0[s], &g->nyl[s]);
        c.y0 = g->y0[s]; c.NY_local = g->nyl[s];
        if (!rc) rc = plbm_create(&c, &g->member[s]);
        if (!rc) rc = plbm_peer_prepare_local(g->member[s]);
    }
    for (int s = 0; s < ndevices && !rc; ++s) rc = plbm_peer_attach_local(g->member[s], g->member.data());
    if (rc) {
        const std::string keep = plbm_last_error();
        plbm_group_destroy(g);
        return group_fail(keep);
    }
    *out = g;
    return 0;
}

void plbm_group_destroy(plbm_group* g)
{
    if (!g) return;
    for (plbm_ctx* c : g->member) if (c) plbm_sync(c);          // nobody still reads a sibling's memory
    for (plbm_ctx* c : g->member) plbm_destroy(c);
    delete g;
}

int plbm_group_size(const plbm_group* g) { return g ? (int)g->member.size() : 0; }
plbm_ctx* plbm_group_member(plbm_group* g, int slab) { return (g && slab >= 0 && slab < (int)g->member.size()) ? g->member[slab] : nullptr; }

int plbm_group_initialize(plbm_group* g)
{
    if (!g) return group_fail("plbm_group_initialize: null group");
    return for_each_slab(g, [&](int s) { return plbm_initialize(g->member[s]); });
}

int plbm_group_step(plbm_group* g, int nsteps, int want_fields)
{
    if (!g) return group_fail("plbm_group_step: null group");
    return for_each_slab(g, [&](int s) { return plbm_step_peer(g->member[s], nsteps, want_fields, nullptr); });
}

int plbm_group_sync(plbm_group* g)
{
    if (!g) return group_fail("plbm_group_sync: null group");
    return for_each_slab(g, [&](int s) { return plbm_sync(g->member[s]) || plbm_peer_check(g->member[s]); });
}

namespace {
// the rows of slab s inside full-lattice arrays (scalar fields are flat x + NX*y, slabs are contiguous row ranges)
void slab_rows(const plbm_group* g, int s, double* const out[PLBM_NUM_FIELDS], double* mine[PLBM_NUM_FIELDS])
{
    for (int k = 0; k < PLBM_NUM_FIELDS; ++k) mine[k] = out[k] ? out[k] + (size_t)g->y0[s] * g->NX : nullptr;
}
}

int plbm_group_download_fields(plbm_group* g, double* const out[PLBM_NUM_FIELDS])
{
    if (!g || !out) return group_fail("plbm_group_download_fields: null argument");
    return for_each_slab(g, [&](int s) {
        double* mine[PLBM_NUM_FIELDS];
        slab_rows(g, s, out, mine);
        return plbm_download_fields(g->member[s], mine);
    });
}

int plbm_group_fetch_begin(plbm_group* g, double* const out[PLBM_NUM_FIELDS])
{
    if (!g || !out) return group_fail("plbm_group_fetch_begin: null argument");
    return for_each_slab(g, [&](int s) {
        double* mine[PLBM_NUM_FIELDS];
        slab_rows(g, s, out, mine);
        return plbm_fetch_begin(g->member[s], mine);
    });
}

int plbm_group_fetch_wait(plbm_group* g)
{
    if (!g) return group_fail("plbm_group_fetch_wait: null group");
    return for_each_slab(g, [&](int s) { return plbm_fetch_wait(g->member[s]); });
}

} // extern "C"
