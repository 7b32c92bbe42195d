/*
 * oracle/plasma_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("port") of the reference's three-species D2Q9 plasma LBM time step
 * (AMSC-24-25/12-lb-12-lb, /root/reference/src/{plasma,collisions,streaming,poisson}.cpp).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it; the product library never does.
 *
 * Pinned against the UNMODIFIED reference compiled in place (oracle/_ref/ref_plasma_parity,
 * built with the reference flags + -ffp-contract=off): tests/test_oracle_vs_ref.py demands
 * bit-identical fields.  The reference itself ships no tests or golden vectors,
 * and FFTW (its Poisson FFT) is absent, so at the FFT boundary parity is "unpinned" and rests on
 * oracle/fft_oracle.c (cross-checked against numpy.fft).
 *
 * Layout contract = the reference's (include/utils.hpp:6-11): Q-arrays i + 9*(x + NX*y),
 * scalar fields x + NX*y.  Species index 0 = electrons, 1 = ions, 2 = neutrals;
 * pair index 0 = e-i, 1 = e-n, 2 = i-n.
 */
#ifndef PLBM_PLASMA_ORACLE_H
#define PLBM_PLASMA_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { PO_POISSON_NONE = 0, PO_POISSON_GS = 1, PO_POISSON_SOR = 2, PO_POISSON_FFT = 3, PO_POISSON_NPS = 4 };
enum { PO_BC_PERIODIC = 0, PO_BC_BOUNCEBACK = 1 };

/* lattice-unit constants, plasma.hpp:86-133 */
typedef struct po_units {
    double cs2, Kb;
    double Ex_ext, Ey_ext;
    double T_init[3];
    double m[3];
    double q[3];          /* q[2] = 0 (neutrals carry no charge) */
    double rho_init[3];
} po_units;

void po_units_from_si(int Z_ion, int A_ion, double Ex_SI, double Ey_SI,
                      double T_e_SI, double T_i_SI, double T_n_SI,
                      double n_e_SI, double n_n_SI, po_units* out);

typedef struct po_state po_state;

po_state* po_create(int NX, int NY, const po_units* u, int poisson_type, int bc_type, double omega_sor);
void po_destroy(po_state* s);

/* plasma.cpp:131-158 */
void po_initialize(po_state* s);
/* individual phases, in the order of the loop body plasma.cpp:476-513 */
void po_update_macro(po_state* s);          /* plasma.cpp:317-456 */
void po_compute_equilibrium(po_state* s);   /* plasma.cpp:162-308 */
void po_thermal_collisions(po_state* s);    /* collisions.cpp:64-122 */
void po_collisions(po_state* s);            /* collisions.cpp:128-181 */
void po_stream(po_state* s);                /* streaming.cpp:13-30 (both distributions) */
void po_solve_poisson(po_state* s);         /* poisson.cpp:25-82 */
void po_step(po_state* s, int nsteps);

/* raw views (owned by the state).  what: */
enum {
    PO_F = 0, PO_G = 1, PO_TMP = 2,               /* idx = species            (9*NX*NY doubles) */
    PO_FEQ_SELF = 3, PO_GEQ_SELF = 4,             /* idx = species                              */
    PO_FEQ_CROSS = 5, PO_GEQ_CROSS = 6,           /* idx = 2*species + slot (slot 0/1, see .c)  */
    PO_RHO = 7, PO_UX = 8, PO_UY = 9, PO_T = 10,  /* idx = species            (NX*NY doubles)   */
    PO_UX_PAIR = 11, PO_UY_PAIR = 12,             /* idx = pair                                 */
    PO_EX = 13, PO_EY = 14, PO_RHO_Q = 15, PO_PHI = 16
};
double* po_field(po_state* s, int what, int idx);

#ifdef __cplusplus
}
#endif
#endif
