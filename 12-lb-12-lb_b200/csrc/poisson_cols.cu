// poisson_cols.cu -- P2 of the spectral Poisson solve (see poisson_fft.cu): columns forward, symbol, inverse.
#include <cstdlib>
#include "poisson_fft_kernels.cuh"

namespace plbm {

// element i (row index of the whole lattice) of spectral column kl: in T2 = [rank s][k_local][rows of s] after
// an all-to-all, or (peer != nullptr) in place in the owning slab's T1 = [k][rows of s], reached through peer memory
struct ColumnIO {
    cpx* T; const SlabTable* tab; const PeerTable* peer; int kl; int kg; int n0;
    static constexpr bool is_smem = false;
    __device__ __forceinline__ cpx* at(int i) const
    {
        if (tab->nranks == 1) return T + (size_t)kl * n0 + i;
        int sr = 0;
        while (i >= tab->y0[sr + 1]) ++sr;
        const int rows = tab->y0[sr + 1] - tab->y0[sr];
        if (peer) return peer->t1[sr] + (size_t)kg * rows + (i - tab->y0[sr]);
        return T + (size_t)tab->nkl * tab->y0[sr] + (size_t)kl * rows + (i - tab->y0[sr]);
    }
    __device__ __forceinline__ cpx load(int i) const
    {
        const double2 t = *reinterpret_cast<const double2*>(at(i));
        return { t.x, t.y };
    }
    __device__ __forceinline__ void store(int i, cpx v) const { *reinterpret_cast<double2*>(at(i)) = make_double2(v.re, v.im); }
};

// last pass of the forward column transform: phi_hat = rho_hat / denom (poisson.cpp:388-409), left in shared memory
struct SymbolOut {
    cpx* buf; const double* sx2; double syk;
    static constexpr bool is_smem = true;
    __device__ __forceinline__ void store(int i, cpx v) const
    {
        const double denom = __dmul_rn(4.0, __dadd_rn(__ldg(sx2 + i), syk));
        if (denom > 1e-15) {
            v.re = __ddiv_rn(v.re, denom);
            v.im = __ddiv_rn(v.im, denom);
        } else {
            v.re = 0.0; v.im = 0.0;
        }
        buf[fft_slot(i)] = v;
    }
};

// One spectral column (all kx): forward, division by the symbol, inverse -- the column never leaves the SM.
template <int FFT_CAP, int TAIL, int ODD>
__global__ void __launch_bounds__(FFT_CAP, 1)
poisson_cols_kernel(cpx* T, const __grid_constant__ FftPlan plan,
                    const double* __restrict__ sx2, const double* __restrict__ sy2, int n0,
                    const __grid_constant__ SlabTable tab, int k0, const __grid_constant__ PeerTable peer, int use_peer)
{
    extern __shared__ cpx fbuf[];
    const int kl = blockIdx.x;
    const ColumnIO col{ T, &tab, use_peer ? &peer : nullptr, kl, k0 + kl, n0 };
    const SymbolOut div{ fbuf, sx2, __ldg(sy2 + k0 + kl) };
    const FftSmem sm{ fbuf };
    fft_run<-1, TAIL, ODD>(plan, fbuf, col, div);
    fft_run<+1, TAIL, ODD>(plan, fbuf, sm, col);
}


// ---- P2 over peer memory with the gather off the FFT's critical path ------------------------------------------------------------
// poisson_cols_kernel with peers is one CTA per SM at the long lengths (a column is 100-140 KB of shared memory), and all CTAs of
// a wave wait for their remote loads together, compute together and store together: at 8 GPUs the kernel takes the SUM of its
// NVLink time and its arithmetic (296 us at 8192^2 against ~130 us of arithmetic).  Here a few COPIER CTAs (the lowest block
// indices, so they are scheduled first and never wait for anybody) pull the slabs' shares of this rank's columns into the
// contiguous local buffer T2 = [k_local][n0], group after group, and publish a flag per group; every other CTA owns one column (CTAs
// start in index order, the order of the gather), waits for its group's flag, transforms the column out of T2 and writes the result straight into the owning slabs' T1 (posted
// stores).  The transfers then run beside the arithmetic instead of between it.  Same arithmetic per column, bit-identical.
struct GatherArgs {
    cpx* T2;               // [nkl][n0]
    unsigned* flags;       // [2 + g]: epoch at which group g was gathered
    unsigned epoch;        // 1, 2, 3, ... per launch
    int ncopy;             // copier CTAs
    int group;             // columns per group
    int vec32;             // every slab boundary and n0 are even: 256-bit accesses
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// source of element i of spectral column kg in the owning slab's T1
__device__ __forceinline__ const cpx* peer_elem(const SlabTable& tab, const PeerTable& peer, int kg, int i)
{
    int sr = 0;
    while (i >= tab.y0[sr + 1]) ++sr;
    const int rows = tab.y0[sr + 1] - tab.y0[sr];
    return peer.t1[sr] + (size_t)kg * rows + (i - tab.y0[sr]);
}

// ---- the copier: bulk copies by the TMA engine, global (a peer's T1 over NVLink, or the own T1) -> shared -> global (T2) ----------
// One thread drives a ring of NSLOT pieces of the CTA's shared memory (the FFT buffer this CTA does not transform in): loads run LAG
// pieces ahead of the stores, so ~LAG pieces (up to 96 KB) are in flight per copier without a register or an address instruction
// per element.  A group's flag is published once the bulk stores of all its pieces have completed.
// (A copier built on per-thread 256-bit loads, 12 in flight per thread, was measured on 8 B200s as well: slower than this ring --
// 321 vs 239 us at 8192^2 -- and dropped.)
__device__ __forceinline__ unsigned p2_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void p2_mbar_wait(unsigned long long* bar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(p2_smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

struct PieceCursor {          // walks the pieces of this copier's groups: group -> column -> slab -> offset inside the slab's share
    int g, kl, kl_end, sr, off;
    __device__ __forceinline__ void start_group(int group, int nkl) { kl = g * group; kl_end = min(kl + group, nkl); sr = 0; off = 0; }
};

constexpr int P2_NSLOT = 8, P2_LAG = 6, P2_PIECE_MAX = 1024;     // 1024 complex = 16 KB per piece
constexpr int P2_PUBLISH_LAG = 3;                                // a group's flag follows its last store by this many stores

static __device__ __noinline__ void copier_tma(const SlabTable& tab, const PeerTable& peer, const GatherArgs& ga, int n0, int k0, cpx* ring, int piece,
                           unsigned long long* bars)
{
    const int nkl = tab.nkl, R = tab.nranks;
    const int ngroups = (nkl + ga.group - 1) / ga.group;
    unsigned* ready = ga.flags + 2;
    for (int i = 0; i < P2_NSLOT; ++i)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(p2_smem_u32(bars + i)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");

    auto address = [&](const PieceCursor& c, const cpx*& src, cpx*& dst, int& n) {
        const int y0 = tab.y0[c.sr], rows = tab.y0[c.sr + 1] - y0;
        n = min(piece, rows - c.off);
        src = peer.t1[c.sr] + (size_t)(k0 + c.kl) * rows + c.off;
        dst = ga.T2 + (size_t)c.kl * n0 + y0 + c.off;
    };
    // returns the group that ended with the piece just left, or -1
    auto advance = [&](PieceCursor& c) {
        c.off += piece;
        if (c.off < tab.y0[c.sr + 1] - tab.y0[c.sr]) return -1;
        c.off = 0;
        if (++c.sr < R) return -1;
        c.sr = 0;
        if (++c.kl < c.kl_end) return -1;
        const int ended = c.g;
        c.g += ga.ncopy;
        if (c.g < ngroups) c.start_group(ga.group, nkl);
        return ended;
    };

    PieceCursor ld, st;
    ld.g = st.g = blockIdx.x;
    if (ld.g >= ngroups) return;
    ld.start_group(ga.group, nkl);
    st = ld;
    int loads = 0, stores = 0;                 // pieces issued so far
    int pend_g[16] = { 0 }, pend_last[16] = { 0 }, head = 0, tail = 0;      // groups whose stores are all issued, waiting for their completion
    auto publish_through = [&](int last_complete) {
        bool any = false;
        while (head != tail && pend_last[head & 15] <= last_complete) {
            if (!any) { asm volatile("fence.proxy.async;" ::: "memory"); __threadfence(); any = true; }
            st_release_gpu(ready + pend_g[head & 15], ga.epoch);
            ++head;
        }
    };
    auto store_next = [&]() {
        const int slot = stores % P2_NSLOT;
        p2_mbar_wait(bars + slot, (unsigned)(stores / P2_NSLOT) & 1u);
        const cpx* src; cpx* dst; int n;
        address(st, src, dst, n);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     :: "l"(dst), "r"(p2_smem_u32(ring + (size_t)slot * piece)), "r"(n * (int)sizeof(cpx)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        const int ended = advance(st);
        if (ended >= 0) { pend_g[tail & 15] = ended; pend_last[tail & 15] = stores; ++tail; }
        ++stores;
    };
    while (ld.g < ngroups) {
        if (loads >= P2_NSLOT) {
            // the slot's previous piece (store number loads - NSLOT) must have LEFT shared memory: with LAG = NSLOT - 2 that is
            // every store but the latest.  (Waiting for the stores' completion here instead costs a write round trip per piece:
            // 7 GB/s per copier instead of 30, measured.)
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        if (head != tail && pend_last[head & 15] <= stores - 1 - P2_PUBLISH_LAG) {
            // a finished group whose last store is PUBLISH_LAG stores old: by now it has completed, the wait costs nothing
            asm volatile("cp.async.bulk.wait_group %0;" :: "n"(P2_PUBLISH_LAG) : "memory");
            publish_through(stores - 1 - P2_PUBLISH_LAG);
        }
        {
            const int slot = loads % P2_NSLOT;
            const cpx* src; cpx* dst; int n;
            address(ld, src, dst, n);
            const unsigned bytes = (unsigned)n * (unsigned)sizeof(cpx);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(p2_smem_u32(bars + slot)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(p2_smem_u32(ring + (size_t)slot * piece)), "l"(src), "r"(bytes), "r"(p2_smem_u32(bars + slot)) : "memory");
            advance(ld);
            ++loads;
        }
        if (loads - stores > P2_LAG) store_next();
    }
    while (stores < loads) store_next();
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    publish_through(stores - 1);
}

template <int W>   // W = 1: one complex per access, 2: two (256-bit)
static __device__ __noinline__ void gather_group(const SlabTable& tab, const PeerTable& peer, cpx* T2, int n0, int k0, int kl0, int ncols)
{
    constexpr int U = 8;                                          // accesses in flight per thread
    const int per_col = n0 / W;
    const int total = ncols * per_col;
    for (int base = threadIdx.x; base < total; base += U * blockDim.x) {
        cpx a[U][W];
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = base + u * blockDim.x;
            if (e < total) {
                const int c = e / per_col, i = (e - c * per_col) * W;
                const cpx* src = peer_elem(tab, peer, k0 + kl0 + c, i);
                if constexpr (W == 2) load_pair(src, a[u][0], a[u][1]);
                else { const double2 t = *reinterpret_cast<const double2*>(src); a[u][0] = { t.x, t.y }; }
            }
        }
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = base + u * blockDim.x;
            if (e < total) {
                const int c = e / per_col, i = (e - c * per_col) * W;
                cpx* dst = T2 + (size_t)(kl0 + c) * n0 + i;
                if constexpr (W == 2) store_pair(dst, a[u][0], a[u][1]);
                else *reinterpret_cast<double2*>(dst) = make_double2(a[u][0].re, a[u][0].im);
            }
        }
    }
}

// a gathered column: loads from its contiguous copy in T2 (`T` is the column's base there; read past L1: other SMs wrote it
// during this kernel), stores into the owning slabs' T1 as ColumnIO does
struct GatheredColumnIO : ColumnIO {
    __device__ __forceinline__ cpx load(int i) const
    {
        const double2 t = __ldcg(reinterpret_cast<const double2*>(T + i));
        return { t.x, t.y };
    }
};

template <int FFT_CAP, int TAIL, int ODD>
__global__ void __launch_bounds__(FFT_CAP, 1)
poisson_cols_gather_kernel(const __grid_constant__ FftPlan plan, const double* __restrict__ sx2, const double* __restrict__ sy2, int n0,
                           const __grid_constant__ SlabTable tab, int k0, const __grid_constant__ PeerTable peer,
                           const __grid_constant__ GatherArgs ga)
{
    extern __shared__ cpx fbuf[];
    __shared__ __align__(8) unsigned long long p2_bars[P2_NSLOT];
    const int nkl = tab.nkl;
    unsigned* ready = ga.flags + 2;
    if ((int)blockIdx.x < ga.ncopy) {
        // ---- copier ----
        const int ngroups = (nkl + ga.group - 1) / ga.group;
        const int fbuf_elems = fft_smem_elems(n0);
        const int piece = min(P2_PIECE_MAX, fbuf_elems / P2_NSLOT);
        if (piece >= 1) {                                 // bulk copies by the TMA engine, driven by one thread
            if (threadIdx.x == 0) copier_tma(tab, peer, ga, n0, k0, fbuf, piece, p2_bars);
            return;
        }
        for (int g = blockIdx.x; g < ngroups; g += ga.ncopy) {
            const int kl0 = g * ga.group, ncols = min(ga.group, nkl - kl0);
            if (ga.vec32) gather_group<2>(tab, peer, ga.T2, n0, k0, kl0, ncols);
            else gather_group<1>(tab, peer, ga.T2, n0, k0, kl0, ncols);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) st_release_gpu(ready + g, ga.epoch);
        }
        return;
    }
    // ---- transform: CTA ncopy + kl owns column kl (the hardware hands CTAs out in index order, the order of the gather) ----
    const int kl = (int)blockIdx.x - ga.ncopy;
    if (threadIdx.x == 0) {
        const unsigned* f = ready + kl / ga.group;
        unsigned spins = 0;
        while (ld_acquire_gpu(f) != ga.epoch) {
            __nanosleep(64);
            if (++spins > (1u << 25)) break;              // ~seconds: give up rather than hang the device (results are then wrong, the parity gates see it)
        }
    }
    __syncthreads();
    const GatheredColumnIO col{ { ga.T2 + (size_t)kl * n0, &tab, &peer, kl, k0 + kl, n0 } };
    const SymbolOut div{ fbuf, sx2, __ldg(sy2 + k0 + kl) };
    const FftSmem sm{ fbuf };
    fft_run<-1, TAIL, ODD>(plan, fbuf, col, div);
    fft_run<+1, TAIL, ODD>(plan, fbuf, sm, col);
}

cudaError_t configure_poisson_cols(const PoissonFftDev& p)
{
    return with_shape(p.col, [&](auto CAP, auto TAIL, auto ODD) {
        cudaError_t e = allow_smem(poisson_cols_kernel<decltype(CAP)::value, decltype(TAIL)::value, decltype(ODD)::value>);
        if (e != cudaSuccess || p.tab.nranks == 1) return e;
        return allow_smem(poisson_cols_gather_kernel<decltype(CAP)::value, decltype(TAIL)::value, decltype(ODD)::value>);
    });
}

cudaError_t launch_poisson_cols(const PoissonFftDev& p, cudaStream_t stream, const PeerTable* peer)
{
    if (p.tab.nkl <= 0) return cudaSuccess;
    const PeerTable none = {};
    const PeerTable& pt = peer ? *peer : none;
    const int use_peer = peer != nullptr;
    const int t = p.col.threads;
    const size_t sm = fft_smem_bytes(p.n0);
    return with_shape(p.col, [&](auto CAP, auto TAIL, auto ODD) {
        poisson_cols_kernel<decltype(CAP)::value, decltype(TAIL)::value, decltype(ODD)::value>
            <<<p.tab.nkl, t, sm, stream>>>(p.T2, p.col, p.sx2, p.sy2, p.n0, p.tab, p.k0, pt, use_peer);
        return cudaGetLastError();
    });
}


// P2 through peer memory with copier CTAs (poisson_cols_gather_kernel).  p.p2_flags: nkl + 2 zero-initialised words.
cudaError_t launch_poisson_cols_gather(PoissonFftDev& p, cudaStream_t stream, const PeerTable& peer)
{
    if (p.tab.nkl <= 0) return cudaSuccess;
    // (read once; thread-safe: with PLBM_DEVICES=N one host thread per device comes through here)
    struct Setup { int sms, want_copiers; cudaError_t err; };
    static const Setup setup = []() {
        Setup u = { 0, 24, cudaSuccess };
        int dev = 0;
        u.err = cudaGetDevice(&dev);
        if (u.err == cudaSuccess) u.err = cudaDeviceGetAttribute(&u.sms, cudaDevAttrMultiProcessorCount, dev);
        if (const char* env = std::getenv("PLBM_P2_COPIERS")) u.want_copiers = std::atoi(env);
        if (u.want_copiers < 1) u.want_copiers = 1;
        return u;
    }();
    if (setup.err != cudaSuccess || setup.sms < 2) return setup.err != cudaSuccess ? setup.err : cudaErrorInvalidDevice;
    const int sms = setup.sms, want_copiers = setup.want_copiers;
    GatherArgs ga;
    ga.T2 = p.T2;
    ga.flags = p.p2_flags;
    ga.epoch = ++p.p2_epoch;
    ga.vec32 = (p.n0 % 2 == 0);
    for (int r = 0; r <= p.tab.nranks; ++r) if (p.tab.y0[r] % 2) ga.vec32 = 0;
    // ~256 KB per group: long enough for the copier's loads in flight, short enough that the first columns are ready early
    int group = (int)((256 * 1024) / (sizeof(cpx) * (size_t)p.n0));
    ga.group = group < 1 ? 1 : (group > 16 ? 16 : group);
    const int ngroups = (p.tab.nkl + ga.group - 1) / ga.group;
    const int t = p.col.threads;
    const size_t sm = fft_smem_bytes(p.n0);
    const int nfft = p.tab.nkl;
    // Copiers take CTA slots from the transforms (at the long lengths a CTA fills an SM).  As few waves of transforms as possible
    // first -- 513 columns on 128 SMs are five waves, on 129 four -- then as many copiers as that leaves, up to `want_copiers`.
    int per_sm = p.p2_per_sm;
    if (per_sm < 1) with_shape(p.col, [&](auto CAP, auto TAIL, auto ODD) {
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, poisson_cols_gather_kernel<decltype(CAP)::value, decltype(TAIL)::value, decltype(ODD)::value>, t, sm);
    });
    if (per_sm < 1) per_sm = 1;
    p.p2_per_sm = per_sm;                                 // (queried once per plan)
    const int slots = sms * per_sm;
    auto waves = [&](int c) { return (nfft + (slots - c) - 1) / (slots - c); };
    int lo = 8 < ngroups ? 8 : ngroups;
    if (lo > slots / 2) lo = slots / 2;
    if (lo < 1) lo = 1;
    int hi = want_copiers < ngroups ? want_copiers : ngroups;
    if (hi > slots / 2) hi = slots / 2;
    if (hi < lo) hi = lo;
    ga.ncopy = lo;
    for (int c = lo; c <= hi; ++c) if (waves(c) <= waves(lo)) ga.ncopy = c;
    return with_shape(p.col, [&](auto CAP, auto TAIL, auto ODD) {
        poisson_cols_gather_kernel<decltype(CAP)::value, decltype(TAIL)::value, decltype(ODD)::value>
            <<<ga.ncopy + nfft, t, sm, stream>>>(p.col, p.sx2, p.sy2, p.n0, p.tab, p.k0, peer, ga);
        return cudaGetLastError();
    });
}

} // namespace plbm
