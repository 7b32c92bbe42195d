/*
 * plbm.h -- C ABI of the B200-native plasma lattice-Boltzmann step (libplbm.so).
 *
 * This is the drop-in boundary for ONE path of AMSC-24-25/12-lb-12-lb: the loop body of
 * LBmethod::Run_simulation (reference src/plasma.cpp:476-523).  The reference has no FFI of its
 * own (it is a single C++ program), so the boundary sits where a maintainer would cut it: under
 * the public C++ headers.  include/plasma.hpp, collisions.hpp, streaming.hpp and poisson.hpp of
 * this repo keep the reference's signatures and forward to the functions below; INTEGRATION.md
 * shows the binding.  Plain pointers and sizes only; every function returns 0 on success and a
 * non-zero code otherwise, with the text available from plbm_last_error().  There is no CPU
 * fallback: without a CUDA device plbm_create fails.
 *
 * Layout contract (reference include/utils.hpp:6-11): population ("Q") arrays are
 * i + 9*(x + NX*y), scalar fields x + NX*y, FP64 throughout.  Species order everywhere:
 * 0 electrons, 1 ions, 2 neutrals.
 */
#ifndef PLBM_H
#define PLBM_H

#ifdef __cplusplus
extern "C" {
#endif

/* enum values of poisson::PoissonType (reference include/poisson.hpp:15-21) */
enum { PLBM_POISSON_NONE = 0, PLBM_POISSON_GS = 1, PLBM_POISSON_SOR = 2, PLBM_POISSON_FFT = 3, PLBM_POISSON_NPS = 4 };
/* enum values of streaming::BCType (reference include/streaming.hpp:10-13) */
enum { PLBM_BC_PERIODIC = 0, PLBM_BC_BOUNCEBACK = 1 };

/* fields handed to visualize::UpdateVisualization (reference include/visualize.hpp:53-61), in its
 * parameter order, followed by the potential */
enum {
    PLBM_F_UX_E = 0, PLBM_F_UY_E, PLBM_F_UX_I, PLBM_F_UY_I, PLBM_F_UX_N, PLBM_F_UY_N,
    PLBM_F_T_E, PLBM_F_T_I, PLBM_F_T_N, PLBM_F_RHO_E, PLBM_F_RHO_I, PLBM_F_RHO_N,
    PLBM_F_RHO_Q, PLBM_F_EX, PLBM_F_EY, PLBM_F_PHI, PLBM_NUM_FIELDS
};

typedef struct plbm_ctx plbm_ctx;

/* Replaces the constructor arguments and in-class initialisers of LBmethod
 * (reference include/plasma.hpp:34-49, 86-133). */
typedef struct plbm_config {
    int NX, NY;               /* global lattice */
    int poisson_type;         /* PLBM_POISSON_* */
    int bc_type;              /* PLBM_BC_* */
    double omega_sor;
    /* lattice-unit constants; fill with plbm_units_from_si() */
    double cs2, Kb;
    double Ex_ext, Ey_ext;
    double T_init[3];
    double m[3];
    double q[3];
    double rho_init[3];
    /* y-slab decomposition: this process owns rows [y0, y0 + NY_local) of the global lattice.
     * nranks == 1 means the whole lattice (y0 = 0, NY_local = NY). */
    int rank, nranks;
    int y0, NY_local;
    int device;               /* CUDA device ordinal, -1 = current device */
    int fields_only;          /* 1: no populations (a context that only serves plbm_host_solve_poisson & co.) */
} plbm_config;

const char* plbm_last_error(void);

/* SI -> lattice units in the expression order of reference include/plasma.hpp:76-133.
 * Fills cs2, Kb, E*_ext, T_init, m, q, rho_init of *cfg; leaves the other members alone. */
int plbm_units_from_si(int Z_ion, int A_ion, double Ex_SI, double Ey_SI,
                       double T_e_SI, double T_i_SI, double T_n_SI,
                       double n_e_SI, double n_n_SI, plbm_config* cfg);

/* LBmethod::LBmethod allocation part (reference src/plasma.cpp:58-120): device state, E = E_ext. */
int plbm_create(const plbm_config* cfg, plbm_ctx** out);
void plbm_destroy(plbm_ctx* ctx);

/* LBmethod::Initialize (reference src/plasma.cpp:131-158), evaluated on the device. */
int plbm_initialize(plbm_ctx* ctx);

/* Populations at the top of the time loop, host AoS arrays of the LOCAL slab
 * (9*NX*NY_local doubles each): f[3], g[3]. */
int plbm_upload_state(plbm_ctx* ctx, const double* const f[3], const double* const g[3]);
int plbm_download_state(plbm_ctx* ctx, double* const f[3], double* const g[3]);
/* Electric field used by the next step (NX*NY_local doubles each). */
int plbm_set_efield(plbm_ctx* ctx, const double* Ex, const double* Ey);

/* nsteps iterations of the loop body of LBmethod::Run_simulation (reference src/plasma.cpp:476-513):
 * UpdateMacro, ComputeEquilibrium, Collide, Stream, SolvePoisson.  With want_fields != 0 the last
 * step also stores the 12 moment fields, so that plbm_download_fields() returns exactly what the
 * reference passes to visualize::UpdateVisualization for that step.  Asynchronous. */
int plbm_step(plbm_ctx* ctx, int nsteps, int want_fields);
int plbm_sync(plbm_ctx* ctx);

/* Copies the requested fields of the local slab to host memory (NX*NY_local doubles each);
 * out[k] == NULL skips field k.  Synchronises. */
int plbm_download_fields(plbm_ctx* ctx, double* const out[PLBM_NUM_FIELDS]);

/* Pipelined variant of plbm_download_fields for loops that hand every step's fields to a consumer
 * (LBmethod::Run_simulation -> visualize::UpdateVisualization): plbm_fetch_begin starts copying the fields of
 * the LAST step into out[] on a separate copy stream and returns at once; the next plbm_step may be issued
 * immediately and overlaps the transfer (the library keeps the data of the pending fetch intact: alternating
 * moment buffers, device snapshots of rho_q / E / phi).  plbm_fetch_wait blocks until out[] is complete.
 * out[] should be page-locked (cudaHostRegister / pinned) for the copy to be asynchronous.  At most one fetch
 * may be pending. */
int plbm_fetch_begin(plbm_ctx* ctx, double* const out[PLBM_NUM_FIELDS]);
int plbm_fetch_wait(plbm_ctx* ctx);
/* Alternate output path.  visualize::UpdateVisualization narrows 12 quantities to CV_32F matrices
 * (reference src/visualize.cpp:226-315: rho_e, rho_i, rho_q, ux_e, uy_e, |u_e|, ux_i, uy_i, |u_i|, T_e, T_i, T_n, each
 * mat(y, x) = float(field[x + NX*y])) and records 19 series at 9 sample points (:168-219: ux, uy, |u| of e, i, n; T of
 * e, i, n; rho of e, i, n; rho_q, Ex, Ey, |E|).  plbm_frames_begin forms exactly these on the device after a
 * plbm_step(..., want_fields = 1) and starts copying them: 48 B per cell instead of 120 cross PCIe, bit-identical
 * to what the unchanged visualiser would have built from the FP64 fields.  frames[k] = NX*NY_local floats (NULL:
 * skip), series = [19][9] doubles (NULL: skip; on a slab, points outside it are left untouched).  Complete it with
 * plbm_fetch_wait; the same one-pending rule as plbm_fetch_begin applies, and the next plbm_step may be issued
 * at once.  A visualiser consumes them through visualize::UpdateVisualizationFrames (include/visualize_frames.hpp). */
#define PLBM_NUM_FRAMES 12
#define PLBM_NUM_SERIES 19
#define PLBM_NUM_POINTS 9
int plbm_frames_begin(plbm_ctx* ctx, float* const frames[PLBM_NUM_FRAMES], double* series);

/* Page-lock / release a host range owned by the caller (cudaHostRegister) so that plbm_fetch_begin, plbm_upload_state
 * and plbm_download_* move it at full PCIe rate and asynchronously.  Optional: unpinned memory stays correct. */
int plbm_pin_host(void* p, size_t bytes);
int plbm_unpin_host(void* p);

/* Same as plbm_step but timed with CUDA events on the library's stream.  Any of the outputs may
 * be NULL.  ms_k1 / ms_poisson are the summed durations of the fused collide-stream kernel and of
 * the Poisson + field kernels; launches = number of kernels launched. */
int plbm_step_timed(plbm_ctx* ctx, int nsteps, int want_fields, float* ms_total, float* ms_k1, float* ms_poisson,
                    long long* launches);

/* ---------------------------------------------------------------------------------------------
 * Stateless per-phase entry points on HOST arrays in the reference layout: upload -> kernel ->
 * download.  They back the free functions of include/collisions.hpp / streaming.hpp (semantic
 * drop-in, not the fast path) and let tests compare single phases.  Arrays of 3 are per species;
 * arrays of 9 equilibria are ordered  e, i, n, e_i, e_n, i_n, i_e, n_e, n_i  exactly like the
 * parameter lists of the reference (include/collisions.hpp:10-29).
 * ------------------------------------------------------------------------------------------- */
/* LBmethod::UpdateMacro (reference src/plasma.cpp:317-456); upx/upy = u_e_i, u_e_n, u_i_n */
int plbm_host_update_macro(int NX, int NY, const plbm_config* units,
                           const double* const f[3], const double* const g[3], const double* Ex, const double* Ey,
                           double* const rho[3], double* const ux[3], double* const uy[3], double* const T[3],
                           double* const upx[3], double* const upy[3], double* rho_q);
/* LBmethod::ComputeEquilibrium (reference src/plasma.cpp:162-308) */
int plbm_host_equilibrium(int NX, int NY, const plbm_config* units,
                          const double* const rho[3], const double* const ux[3], const double* const uy[3], const double* const T[3],
                          const double* const upx[3], const double* const upy[3],
                          double* const f_eq[9], double* const g_eq[9]);
/* collisions::ThermalCollisions (reference src/collisions.cpp:64-122): out[3] receives g + C_T + DeltaT */
int plbm_host_thermal_collisions(int NX, int NY, const plbm_config* units,
                                 const double* const g[3], const double* const g_eq[9], const double* const f_eq[9],
                                 const double* const rho[3], const double* const ux[3], const double* const uy[3],
                                 double* const out[3]);
/* collisions::Collisions (reference src/collisions.cpp:128-181): out[3] receives f + C (+ F) */
int plbm_host_collisions(int NX, int NY, const plbm_config* units,
                         const double* const f[3], const double* const f_eq[9],
                         const double* const rho[3], const double* const ux[3], const double* const uy[3],
                         const double* Ex, const double* Ey, double* const out[3]);
/* streaming::StreamingPeriodic / ThermalStreamingPeriodic (reference src/streaming.cpp:35-59): out[3] receives the
 * streamed populations.  bc_type other than PLBM_BC_PERIODIC: see plbm_last_error(). */
int plbm_host_stream(int NX, int NY, int bc_type, const double* const in[3], double* const out[3]);
/* poisson::SolvePoisson (reference src/poisson.cpp:25-82) on a context created for the lattice: rho_q in,
 * Ex/Ey updated in place, potential kept inside the context (warm start, like the reference's static phi). */
int plbm_host_solve_poisson(plbm_ctx* ctx, const double* rho_q, double* Ex, double* Ey);
/* the two halves separately: poisson::SolvePoisson_{GS,SOR,FFT,9point} (rho_q -> phi, kept in the context) and
 * poisson::ComputeElectricField[_Periodic] (phi -> Ex, Ey; Ex/Ey are in-out because walls copy onto the rim) */
int plbm_host_poisson_solver(plbm_ctx* ctx, int poisson_type, const double* rho_q);
/* The reference passes type, boundary and omega to poisson::SolvePoisson on EVERY call (src/poisson.cpp:25-82; only phi's size and
 * the NONE zero-fill are call_once): a fields-only context takes the values for its next plbm_host_* calls from here.  omega_sor is
 * also what plbm_host_poisson_solver(PLBM_POISSON_SOR[_PERIODIC]) uses.  The spectral plan is built on the first FFT request. */
int plbm_host_poisson_config(plbm_ctx* ctx, int poisson_type, int bc_type, double omega_sor);
/* additional solver codes for plbm_host_poisson_solver only: poisson::SolvePoisson_{GS,SOR,9point}_Periodic
 * (reference src/poisson.cpp:146-211, 283-354, 487-546; public in include/poisson.hpp:55-97, not called by the reference's own loop) */
#define PLBM_POISSON_GS_PERIODIC  5
#define PLBM_POISSON_SOR_PERIODIC 6
#define PLBM_POISSON_NPS_PERIODIC 7
int plbm_host_efield(plbm_ctx* ctx, int bc_type, double* Ex, double* Ey);

/* ---------------------------------------------------------------------------------------------
 * Several slabs (one process per GPU).  The lattice is cut along y; the library owns the rule
 * (plbm_slab_of: even-sized slabs, balanced by work: the rows of the reference's central block, where its
 * initial condition puts the plasma, weigh 1.17; PLBM_SLAB_ALPHA=0 in the environment gives equal slabs).  The data path has two exchanges per step, which the
 * host layer performs between these calls (NCCL send/recv and all-to-all in this repo's driver):
 *   plbm_step_local                      K1 on the slab (halo rows must be current)
 *   plbm_halo_pack -> exchange -> plbm_halo_unpack      18 rows of NX doubles per side
 *   plbm_poisson_stage(0) -> all-to-all T1->T2 -> stage(1) -> all-to-all T2->T1 -> stage(2)
 *   exchange one row of phi per side -> plbm_poisson_stage(3)
 * ------------------------------------------------------------------------------------------- */
typedef struct plbm_exchange {
    int nranks, rank;
    void *halo_send_lo, *halo_send_hi, *halo_recv_lo, *halo_recv_hi;   /* device, halo_count doubles each */
    long long halo_count;
    void *phi_first_row, *phi_last_row, *phi_below, *phi_above;         /* device, phi_count doubles each */
    long long phi_count;
    void *t1, *t2;        /* device, complex<double>: T1 = [NY/2+1][rows of this slab], T2 = [rank s][columns of this rank][rows of s] */
    int slab_y0[17];      /* slab s owns rows [slab_y0[s], slab_y0[s+1]) */
    int slab_k0[17];      /* rank d owns spectral columns [slab_k0[d], slab_k0[d+1]) */
} plbm_exchange;

int plbm_slab_of(int NY, int rank, int nranks, int* y0, int* ny_local);
int plbm_step_local(plbm_ctx* ctx, int want_fields);
int plbm_halo_pack(plbm_ctx* ctx);
int plbm_halo_unpack(plbm_ctx* ctx);
int plbm_poisson_stage(plbm_ctx* ctx, int stage);
int plbm_exchange_info(plbm_ctx* ctx, plbm_exchange* out);

/* Peer-memory transposes (all slabs on GPUs of one node with peer access, one process per GPU).  Instead of the
 * two all-to-alls, the column pass reads each slab's part of its spectral columns directly from that slab's T1
 * over NVLink and writes the result back in place:
 *   plbm_poisson_stage(0) -> plbm_peer_barrier -> plbm_poisson_stage(4) -> plbm_peer_barrier -> plbm_poisson_stage(2)
 * Setup: every rank calls plbm_peer_export (a PLBM_PEER_BLOB_BYTES blob holding CUDA IPC handles), the host layer
 * gathers the blobs of all ranks in rank order and hands them to plbm_peer_attach.  plbm_peer_barrier enqueues a
 * barrier among all slabs' streams (flags in peer memory; it gives up after ~10 s and plbm_peer_check reports it).
 * When attaching fails (no IPC / no peer access) the host layer keeps using the all-to-all path. */
#define PLBM_PEER_BLOB_BYTES 512
int plbm_peer_export(plbm_ctx* ctx, void* blob);
int plbm_peer_attach(plbm_ctx* ctx, const void* blobs_of_all_ranks);
int plbm_peer_barrier(plbm_ctx* ctx);
int plbm_peer_check(plbm_ctx* ctx);
/* With peers attached the other two exchanges need no messages either: plbm_halo_push packs the outgoing
 * population rows straight into the periodic neighbours' receive buffers (then barrier, plbm_halo_unpack), and
 * plbm_phi_rows_push copies the slab's first/last row of phi into the neighbours' phi_above/phi_below (then
 * barrier, plbm_poisson_stage(3)).  A whole step is then
 *   step_local, halo_push, stage(0), BARRIER, halo_unpack, stage(4), BARRIER, stage(5), BARRIER, stage(3)
 * where stage(5) = stage(2) whose last pass also stores the boundary rows into the neighbours (plbm_phi_rows_push is
 * the same transfer as two copies, for callers that keep stage(2)).
 * Hazards across steps: a rank passes barrier n only after every rank has enqueued-and-finished everything before
 * its own barrier n.  Writes into a peer (halo rows, phi rows, T1 columns in stage 4) always follow a barrier that the
 * peer reaches after its last read of the previous contents: halo buffers are read by halo_unpack (before the 2nd
 * barrier of a step) and rewritten after the 3rd; T1 is read by the peers' stage 4 (before the 2nd barrier) and
 * rewritten by stage 0 of the next step; phi copies are read by K1 / stage 3 (before the next step's 1st barrier) and
 * rewritten after its 2nd. */
int plbm_halo_push(plbm_ctx* ctx);
int plbm_phi_rows_push(plbm_ctx* ctx);
/* Unmap the peers' memory.  Every rank must have detached (host-level barrier) before any rank destroys its context. */
int plbm_peer_detach(plbm_ctx* ctx);
/* A whole time step of a slab with peers attached, nsteps times, in ONE call (no host round trip per kernel): the one-stream
 * sequence documented above.  With PLBM_PEER_PIPELINE=1 in the environment the Poisson solve of step t runs on a second stream
 * BESIDE K1 of step t instead: rho_q of step t is pulled from the planes step t-1 wrote (it does not depend on step t's
 * collision), phi of step t lands in a second potential, and K1 of step t+1 waits for it (bit-identical; measured slower).  stage_ms, if not NULL, must hold
 * PLBM_PEER_STAGES floats and receives the accumulated device time [ms] per stage over at most the last 32 steps:
 *   0 K1   1 halo push   2 halo barrier   3 halo unpack   4 wait for phi (what is left of the solve on the critical path)
 *   5 charge pull   6 P1   7 barrier   8 P2 over peer memory   9 barrier   10 P3 + phi rows   11 barrier
 * (sequential sequence: 2, 4 and 5 stay 0; barrier 7 also orders the halo exchange). */
#define PLBM_PEER_STAGES 12
int plbm_step_peer(plbm_ctx* ctx, int nsteps, int want_fields, float* stage_ms);
/* Slabs that live in one process: wire the peers from the sibling contexts themselves (cudaDeviceEnablePeerAccess, no IPC).
 * plbm_peer_prepare_local on every slab first, then plbm_peer_attach_local(ctx, all) with all[s] = context of slab s. */
int plbm_peer_prepare_local(plbm_ctx* ctx);
int plbm_peer_attach_local(plbm_ctx* ctx, plbm_ctx* const* all);

/* ---------------------------------------------------------------------------------------------
 * Several GPUs behind ONE host object (the reference's host code is one C++ program: LBmethod with PLBM_DEVICES=N uses this).
 * A group owns one context per device (slab s on device first_device + s), wires them through peer memory and drives each
 * from its own host thread; the data path between the slabs is the kernels' own (NVLink), the host only starts the step
 * sequences.  Field arrays are full-lattice host arrays [NY][NX]; every slab copies its own rows.  Periodic boundaries with the
 * spectral (or no) Poisson solve, as for any multi-slab run.
 * ------------------------------------------------------------------------------------------- */
typedef struct plbm_group plbm_group;
int plbm_group_create(const plbm_config* cfg /* global lattice; cfg->device = first device or -1 for 0 */, int ndevices, plbm_group** out);
void plbm_group_destroy(plbm_group* g);
int plbm_group_size(const plbm_group* g);
plbm_ctx* plbm_group_member(plbm_group* g, int slab);
int plbm_group_initialize(plbm_group* g);
int plbm_group_step(plbm_group* g, int nsteps, int want_fields);
int plbm_group_sync(plbm_group* g);
int plbm_group_download_fields(plbm_group* g, double* const out[PLBM_NUM_FIELDS]);
int plbm_group_fetch_begin(plbm_group* g, double* const out[PLBM_NUM_FIELDS]);
int plbm_group_fetch_wait(plbm_group* g);

/* Self-test hook: the library's fast division primitives on caller-supplied operands.
 * mode 0: a[i] / b[i] (data-dependent divisor); mode 1: a[i] / b[0] through the two-term 1/b[0] product (b[0] in {3,5,6});
 * mode 2: a[i] / b[0] through the device-refined reciprocal.  ok[i] = 1 when the operands were inside the domain where the
 * kernel accepts the fast result (otherwise the kernels recompute with IEEE division). */
int plbm_selftest_division(const double* a, const double* b, int n, int mode, double* out, int* ok);

/* Introspection used by the benchmarks and tests. */
int plbm_local_rows(const plbm_ctx* ctx, int* y0, int* ny_local);
long long plbm_device_bytes(const plbm_ctx* ctx);
void* plbm_stream(plbm_ctx* ctx);               /* cudaStream_t the library launches on */

#ifdef __cplusplus
}
#endif
#endif
