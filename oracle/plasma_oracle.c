/*
 * oracle/plasma_oracle.c -- TEST INFRASTRUCTURE ONLY (see plasma_oracle.h).
 *
 * CPU restatement of the reference time step.  Every floating-point expression keeps the
 * reference's evaluation order (C left-to-right association, true IEEE division, no FMA:
 * build with -ffp-contract=off), so the results are bit-identical to the reference compiled
 * with the same contraction setting.  Citations are file:line under /root/reference.
 */
#include "plasma_oracle.h"
#include "fft_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NQ 9

/* D2Q9 tables, src/plasma.cpp:10-16; opposite directions, src/streaming.cpp:8 */
static const int CX[NQ] = { 0, 1, 0, -1, 0, 1, -1, -1, 1 };
static const int CY[NQ] = { 0, 0, 1, 0, -1, 1, 1, -1, -1 };
static const int OPP[NQ] = { 0, 3, 4, 1, 2, 7, 8, 5, 6 };
static double WGT[NQ];

/* relaxation times, src/collisions.cpp:6-7 */
static const double TAU_SELF[3] = { 5.0, 3.0, 1.0 };   /* tau_e, tau_i, tau_n          */
static const double TAU_PAIR[3] = { 6.0, 4.0, 2.0 };   /* tau_e_i, tau_e_n, tau_i_n    */
/* pair used by the two cross terms of each species, in the reference's summation order:
 * e: (e_i, e_n)   i: (i_e, i_n)   n: (n_e, n_i)      collisions.cpp:107-109,166-168 */
static const int PAIR_OF[3][2] = { { 0, 1 }, { 0, 2 }, { 1, 2 } };
/* species forming each pair: ei, en, in                plasma.cpp:426-449 */
static const int PAIR_A[3] = { 0, 0, 1 };
static const int PAIR_B[3] = { 1, 2, 2 };

struct po_state {
    int NX, NY;
    po_units u;
    int poisson_type, bc_type;
    double omega;
    double *f[3], *g[3], *tmp[3];
    double *feq_self[3], *geq_self[3];
    double *feq_cross[3][2], *geq_cross[3][2];
    double *rho[3], *ux[3], *uy[3], *T[3];
    double *uxp[3], *uyp[3];
    double *Ex, *Ey, *rho_q, *phi;
    int poisson_called;
    offt_plan2d* fft;
    double *fft_in, *fft_out;
    offt_cpx *rho_hat, *phi_hat;
};

static inline size_t iq(int x, int y, int i, int NX) { return (size_t)i + NQ * ((size_t)x + (size_t)NX * y); }  /* utils.hpp:6-8 */
static inline size_t ic(int x, int y, int NX) { return (size_t)x + (size_t)NX * y; }                            /* utils.hpp:9-11 */

/* include/plasma.hpp:76-133, same expression order */
void po_units_from_si(int Z_ion, int A_ion, double Ex_SI, double Ey_SI,
                      double T_e_SI, double T_i_SI, double T_n_SI,
                      double n_e_SI, double n_n_SI, po_units* o)
{
    const double kB_SI = 1.380649e-23, e_charge_SI = 1.602176634e-19, epsilon0_SI = 8.854187817e-12;
    const double m_e_SI = 9.10938356e-31, u_SI = 1.66053906660e-27;
    const double m_i_SI = A_ion * u_SI;
    const double m_n_SI = A_ion * u_SI;
    const double n0_SI = n_e_SI;
    const double M0_SI = m_e_SI;
    const double T0_SI = T_e_SI;
    const double Q0_SI = e_charge_SI;
    const double L0_SI = sqrt(epsilon0_SI * kB_SI * T0_SI / (n0_SI * Q0_SI * Q0_SI)) * 1e-2;
    const double t0_SI = sqrt(epsilon0_SI * M0_SI / (3.0 * n0_SI * Q0_SI * Q0_SI)) * 1e-2;
    const double E0_SI = M0_SI * L0_SI / (Q0_SI * t0_SI * t0_SI);
    o->cs2 = kB_SI * T0_SI / M0_SI * t0_SI * t0_SI / (L0_SI * L0_SI);
    o->Kb = kB_SI * (t0_SI * t0_SI * T0_SI) / (L0_SI * L0_SI * M0_SI);
    o->Ex_ext = Ex_SI / E0_SI;
    o->Ey_ext = Ey_SI / E0_SI;
    o->T_init[0] = T_e_SI / T0_SI;
    o->T_init[1] = T_i_SI / T0_SI;
    o->T_init[2] = T_n_SI / T0_SI;
    o->m[0] = m_e_SI / M0_SI;
    o->m[1] = m_i_SI / M0_SI;
    o->m[2] = m_n_SI / M0_SI;
    o->q[0] = -e_charge_SI / Q0_SI;
    o->q[1] = Z_ion * e_charge_SI / Q0_SI;
    o->q[2] = 0.0;
    o->rho_init[0] = o->m[0] * n_e_SI / n0_SI;
    o->rho_init[1] = o->m[1] * n_e_SI / n0_SI / Z_ion;
    o->rho_init[2] = o->m[2] * n_n_SI / n0_SI;
}

static double* dalloc(size_t n, int zero)
{
    double* p = (double*)malloc(sizeof(double) * n);
    if (zero) memset(p, 0, sizeof(double) * n);
    return p;
}

/* LBmethod ctor, plasma.cpp:22-124 (allocation + E = E_ext), without the Initialize() call */
po_state* po_create(int NX, int NY, const po_units* u, int poisson_type, int bc_type, double omega_sor)
{
    WGT[0] = 4.0 / 9.0;
    for (int i = 1; i < 5; ++i) WGT[i] = 1.0 / 9.0;
    for (int i = 5; i < 9; ++i) WGT[i] = 1.0 / 36.0;
    po_state* s = (po_state*)calloc(1, sizeof(*s));
    s->NX = NX; s->NY = NY; s->u = *u;
    s->poisson_type = poisson_type; s->bc_type = bc_type; s->omega = omega_sor;
    const size_t N = (size_t)NX * NY, NQN = N * NQ;
    for (int k = 0; k < 3; ++k) {
        s->f[k] = dalloc(NQN, 1); s->g[k] = dalloc(NQN, 1); s->tmp[k] = dalloc(NQN, 1);
        s->feq_self[k] = dalloc(NQN, 1); s->geq_self[k] = dalloc(NQN, 1);
        for (int c = 0; c < 2; ++c) { s->feq_cross[k][c] = dalloc(NQN, 1); s->geq_cross[k][c] = dalloc(NQN, 1); }
        s->rho[k] = dalloc(N, 1); s->ux[k] = dalloc(N, 1); s->uy[k] = dalloc(N, 1); s->T[k] = dalloc(N, 1);
        s->uxp[k] = dalloc(N, 1); s->uyp[k] = dalloc(N, 1);
    }
    s->Ex = dalloc(N, 0); s->Ey = dalloc(N, 0); s->rho_q = dalloc(N, 1); s->phi = dalloc(N, 1);
    for (size_t c = 0; c < N; ++c) { s->Ex[c] = u->Ex_ext; s->Ey[c] = u->Ey_ext; }   /* plasma.cpp:116-117 */
    return s;
}

void po_destroy(po_state* s)
{
    if (!s) return;
    for (int k = 0; k < 3; ++k) {
        free(s->f[k]); free(s->g[k]); free(s->tmp[k]); free(s->feq_self[k]); free(s->geq_self[k]);
        for (int c = 0; c < 2; ++c) { free(s->feq_cross[k][c]); free(s->geq_cross[k][c]); }
        free(s->rho[k]); free(s->ux[k]); free(s->uy[k]); free(s->T[k]); free(s->uxp[k]); free(s->uyp[k]);
    }
    free(s->Ex); free(s->Ey); free(s->rho_q); free(s->phi);
    if (s->fft) { offt_plan2d_destroy(s->fft); free(s->fft_in); free(s->fft_out); free(s->rho_hat); free(s->phi_hat); }
    free(s);
}

double* po_field(po_state* s, int what, int idx)
{
    switch (what) {
    case PO_F: return s->f[idx];
    case PO_G: return s->g[idx];
    case PO_TMP: return s->tmp[idx];
    case PO_FEQ_SELF: return s->feq_self[idx];
    case PO_GEQ_SELF: return s->geq_self[idx];
    case PO_FEQ_CROSS: return s->feq_cross[idx / 2][idx % 2];
    case PO_GEQ_CROSS: return s->geq_cross[idx / 2][idx % 2];
    case PO_RHO: return s->rho[idx];
    case PO_UX: return s->ux[idx];
    case PO_UY: return s->uy[idx];
    case PO_T: return s->T[idx];
    case PO_UX_PAIR: return s->uxp[idx];
    case PO_UY_PAIR: return s->uyp[idx];
    case PO_EX: return s->Ex;
    case PO_EY: return s->Ey;
    case PO_RHO_Q: return s->rho_q;
    case PO_PHI: return s->phi;
    default: return NULL;
    }
}

/* plasma.cpp:131-158: charged species only in the open central square, neutrals everywhere */
void po_initialize(po_state* s)
{
    const int NX = s->NX, NY = s->NY;
    for (int y = 0; y < NY; ++y)
        for (int x = 0; x < NX; ++x) {
            const int inside = (x >= NX / 4 + 1) && (x < 3 * NX / 4) && (y >= NY / 4 + 1) && (y < 3 * NY / 4);
            for (int i = 0; i < NQ; ++i) {
                const size_t k = iq(x, y, i, NX);
                if (inside) {
                    s->f[0][k] = WGT[i] * s->u.rho_init[0]; s->g[0][k] = WGT[i] * s->u.T_init[0];
                    s->f[1][k] = WGT[i] * s->u.rho_init[1]; s->g[1][k] = WGT[i] * s->u.T_init[1];
                }
                s->f[2][k] = WGT[i] * s->u.rho_init[2];
                s->g[2][k] = WGT[i] * s->u.T_init[2];
            }
        }
}

/* plasma.cpp:317-456 */
void po_update_macro(po_state* s)
{
    const int NX = s->NX, NY = s->NY;
    #pragma omp parallel for schedule(static)
    for (int y = 0; y < NY; ++y)
        for (int x = 0; x < NX; ++x) {
            const size_t c = ic(x, y, NX);
            double rl[3], mx[3], my[3], tl[3];
            for (int k = 0; k < 3; ++k) {
                double r = 0.0, ax = 0.0, ay = 0.0, t = 0.0;
                for (int i = 0; i < NQ; ++i) {                          /* :352-372 */
                    const double fi = s->f[k][iq(x, y, i, NX)];
                    r += fi;
                    ax += fi * CX[i];
                    ay += fi * CY[i];
                    t += s->g[k][iq(x, y, i, NX)];
                }
                rl[k] = r; mx[k] = ax; my[k] = ay; tl[k] = t;
            }
            for (int k = 0; k < 3; ++k) {
                if (rl[k] < 1e-10) {                                    /* :373-377 */
                    s->rho[k][c] = 0.0; s->ux[k][c] = 0.0; s->uy[k][c] = 0.0; s->T[k][c] = 0.0;
                } else if (k < 2) {                                     /* charged: :378-412 */
                    s->rho[k][c] = rl[k];
                    double vx = (mx[k] == rl[k] || mx[k] == -rl[k]) ? 0.0 : mx[k] / rl[k];
                    double vy = (my[k] == rl[k] || my[k] == -rl[k]) ? 0.0 : my[k] / rl[k];
                    vx += 0.5 * s->u.q[k] * s->Ex[c] / s->u.m[k];
                    vy += 0.5 * s->u.q[k] * s->Ey[c] / s->u.m[k];
                    s->ux[k][c] = vx; s->uy[k][c] = vy;
                    s->T[k][c] = tl[k];
                } else {                                                /* neutrals: :419-425 */
                    s->rho[k][c] = rl[k];
                    s->ux[k][c] = mx[k] / rl[k];
                    s->uy[k][c] = my[k] / rl[k];
                    s->T[k][c] = tl[k];
                }
            }
            for (int p = 0; p < 3; ++p) {                               /* :426-449 */
                const int a = PAIR_A[p], b = PAIR_B[p];
                if (rl[a] < 1e-10 && rl[b] < 1e-10) {
                    s->uxp[p][c] = 0.0; s->uyp[p][c] = 0.0;
                } else {
                    s->uxp[p][c] = (rl[a] * s->ux[a][c] + rl[b] * s->ux[b][c]) / (rl[a] + rl[b]);
                    s->uyp[p][c] = (rl[a] * s->uy[a][c] + rl[b] * s->uy[b][c]) / (rl[a] + rl[b]);
                }
            }
            double rq = (s->u.q[1] * s->rho[1][c] / s->u.m[1] + s->u.q[0] * s->rho[0][c] / s->u.m[0]);   /* :452 */
            if (rq < 1e-15) rq = 0.0;                                   /* :453 */
            s->rho_q[c] = rq;
        }
}

/* the bracket of plasma.cpp:195-200 */
static inline double eq_bracket(double cu, double u2, double invcs2)
{
    return 1.0 + cu * invcs2 + (cu * cu) * 0.5 * invcs2 * invcs2 - u2 * 0.5 * invcs2;
}

/* plasma.cpp:162-308 */
void po_compute_equilibrium(po_state* s)
{
    const int NX = s->NX, NY = s->NY;
    const double invcs2 = 1.0 / s->u.cs2;                               /* :164 */
    #pragma omp parallel for schedule(static)
    for (int y = 0; y < NY; ++y)
        for (int x = 0; x < NX; ++x) {
            const size_t c = ic(x, y, NX);
            double u2s[3], u2p[3];
            for (int k = 0; k < 3; ++k) {
                u2s[k] = s->ux[k][c] * s->ux[k][c] + s->uy[k][c] * s->uy[k][c];
                u2p[k] = s->uxp[k][c] * s->uxp[k][c] + s->uyp[k][c] * s->uyp[k][c];
            }
            for (int i = 0; i < NQ; ++i) {
                const size_t q = iq(x, y, i, NX);
                double bs[3], bp[3];
                for (int k = 0; k < 3; ++k) {
                    const double cus = CX[i] * s->ux[k][c] + CY[i] * s->uy[k][c];
                    const double cup = CX[i] * s->uxp[k][c] + CY[i] * s->uyp[k][c];
                    bs[k] = eq_bracket(cus, u2s[k], invcs2);
                    bp[k] = eq_bracket(cup, u2p[k], invcs2);
                }
                for (int k = 0; k < 3; ++k) {
                    s->feq_self[k][q] = WGT[i] * s->rho[k][c] * bs[k];
                    s->geq_self[k][q] = WGT[i] * s->T[k][c] * bs[k];
                    for (int m = 0; m < 2; ++m) {
                        s->feq_cross[k][m][q] = WGT[i] * s->rho[k][c] * bp[PAIR_OF[k][m]];
                        s->geq_cross[k][m][q] = WGT[i] * s->T[k][c] * bp[PAIR_OF[k][m]];
                    }
                }
            }
        }
}

/* one term of collisions.cpp:86-96 */
static inline double thermal_term(double rho, double tau, double feq)
{
    return (2.0 * rho * (1.0 - 1.0 / tau) * (1.0 - 1.0 / tau) - 2.0 * (1.0 - 1.0 / tau) * rho - NQ * feq / tau)
           / (2.0 * (2.0 * (1.0 - 1.0 / tau) + NQ * feq / tau));
}

static void swap3(double** a, double** b)
{
    for (int k = 0; k < 3; ++k) { double* t = a[k]; a[k] = b[k]; b[k] = t; }
}

/* collisions.cpp:64-122 */
void po_thermal_collisions(po_state* s)
{
    const int NX = s->NX, NY = s->NY;
    const double Kb = s->u.Kb;
    #pragma omp parallel for schedule(static)
    for (int y = 0; y < NY; ++y)
        for (int x = 0; x < NX; ++x) {
            const size_t c = ic(x, y, NX);
            for (int i = 0; i < NQ; ++i) {
                const size_t q = iq(x, y, i, NX);
                for (int k = 0; k < 3; ++k) {
                    const double t0 = thermal_term(s->rho[k][c], TAU_SELF[k], s->feq_self[k][q]);
                    const double t1 = thermal_term(s->rho[k][c], TAU_PAIR[PAIR_OF[k][0]], s->feq_cross[k][0][q]);
                    const double t2 = thermal_term(s->rho[k][c], TAU_PAIR[PAIR_OF[k][1]], s->feq_cross[k][1][q]);
                    const double dE = s->rho[k][c] * (t0 + t1 + t2) * (s->ux[k][c] * s->ux[k][c] + s->uy[k][c] * s->uy[k][c]);   /* :98-100 */
                    const double dT = -dE / Kb;                                                                                   /* :102-104 */
                    const double gk = s->g[k][q];
                    const double CT = -(gk - s->geq_self[k][q]) / TAU_SELF[k]
                                      - (gk - s->geq_cross[k][0][q]) / TAU_PAIR[PAIR_OF[k][0]]
                                      - (gk - s->geq_cross[k][1][q]) / TAU_PAIR[PAIR_OF[k][1]];                                    /* :107-109 */
                    s->tmp[k][q] = gk + CT + dT;                                                                                  /* :112-114 */
                }
            }
        }
    swap3(s->g, s->tmp);                                                                                                          /* :119-121 */
}

/* collisions.cpp:128-181 */
void po_collisions(po_state* s)
{
    const int NX = s->NX, NY = s->NY;
    const double cs2 = s->u.cs2;
    #pragma omp parallel for schedule(static)
    for (int y = 0; y < NY; ++y)
        for (int x = 0; x < NX; ++x) {
            const size_t c = ic(x, y, NX);
            const double Ex = s->Ex[c], Ey = s->Ey[c];
            for (int i = 0; i < NQ; ++i) {
                const size_t q = iq(x, y, i, NX);
                for (int k = 0; k < 3; ++k) {
                    const double fk = s->f[k][q];
                    const double C = -(fk - s->feq_self[k][q]) / TAU_SELF[k]
                                     - (fk - s->feq_cross[k][0][q]) / TAU_PAIR[PAIR_OF[k][0]]
                                     - (fk - s->feq_cross[k][1][q]) / TAU_PAIR[PAIR_OF[k][1]];                                     /* :166-168 */
                    if (k < 2) {
                        const double ux = s->ux[k][c], uy = s->uy[k][c];
                        const double F = WGT[i] * s->u.q[k] * s->rho[k][c] / s->u.m[k] / cs2 * (1.0 - 1.0 / (2 * TAU_SELF[k])) * (
                            (CX[i] * Ex + CY[i] * Ey) +
                            (CX[i] * ux + CY[i] * uy) * (CX[i] * Ex + CY[i] * Ey) / cs2 -
                            (ux * Ex + uy * Ey));                                                                                 /* :154-163 */
                        s->tmp[k][q] = fk + C + F;                                                                                /* :171-172 */
                    } else {
                        s->tmp[k][q] = fk + C;                                                                                    /* :173 */
                    }
                }
            }
        }
    swap3(s->f, s->tmp);                                                                                                          /* :178-180 */
}

/* streaming.cpp:35-59 / :117-141 : push with periodic wrap into tmp, then swap */
static void stream_periodic(po_state* s, double** d)
{
    const int NX = s->NX, NY = s->NY;
    #pragma omp parallel for schedule(static)
    for (int y = 0; y < NY; ++y)
        for (int x = 0; x < NX; ++x)
            for (int i = 0; i < NQ; ++i) {
                const int xs = (x + NX + CX[i]) % NX, ys = (y + NY + CY[i]) % NY;
                const size_t to = iq(xs, ys, i, NX), from = iq(x, y, i, NX);
                for (int k = 0; k < 3; ++k) s->tmp[k][to] = d[k][from];
            }
    swap3(d, s->tmp);
}

/* streaming.cpp:66-112 / :150-196 : the reference's (non-standard) bounce-back.  Its loop is an
 * orphaned `omp for`, i.e. it runs serially in x-outer / y / i order, and several destinations
 * collide or are never written (stale tmp shows through); the serial order is kept here. */
static void stream_bounceback(po_state* s, double** d)
{
    const int NX = s->NX, NY = s->NY;
    for (int x = 0; x < NX; ++x)
        for (int y = 0; y < NY; ++y)
            for (int i = 0; i < NQ; ++i) {
                const int xs = x + CX[i], ys = y + CY[i];
                const size_t from = iq(x, y, i, NX);
                size_t to;
                if (xs >= 0 && xs < NX && ys >= 0 && ys < NY) to = iq(xs, ys, i, NX);
                else if (xs >= 0 && xs < NX) to = iq(xs, y, OPP[i], NX);
                else if (ys >= 0 && ys < NY) to = iq(x, ys, OPP[i], NX);
                else to = iq(x, y, OPP[i], NX);
                for (int k = 0; k < 3; ++k) s->tmp[k][to] = d[k][from];
            }
    swap3(d, s->tmp);
}

/* streaming.cpp:13-30 */
void po_stream(po_state* s)
{
    if (s->bc_type == PO_BC_PERIODIC) { stream_periodic(s, s->f); stream_periodic(s, s->g); }
    else { stream_bounceback(s, s->f); stream_bounceback(s, s->g); }
}

/* poisson.cpp:365-420 + 611-623; note (n0,n1) = (NX,NY) on x-fastest data (a14 quirk) */
static void poisson_fft(po_state* s)
{
    const int NX = s->NX, NY = s->NY, NYh = NY / 2 + 1;
    const size_t N = (size_t)NX * NY;
    if (!s->fft) {
        s->fft = offt_plan2d_create(NX, NY);
        s->fft_in = dalloc(N, 1); s->fft_out = dalloc(N, 1);
        s->rho_hat = (offt_cpx*)malloc(sizeof(offt_cpx) * (size_t)NX * NYh);
        s->phi_hat = (offt_cpx*)malloc(sizeof(offt_cpx) * (size_t)NX * NYh);
    }
    for (size_t c = 0; c < N; ++c) s->fft_in[c] = s->rho_q[c];
    offt_r2c_2d(s->fft, s->fft_in, s->rho_hat);
    for (int i = 0; i < NX; ++i)
        for (int j = 0; j < NYh; ++j) {
            const int kx = (i <= NX / 2) ? i : i - NX;
            const int ky = j;
            const double sinx = sin(M_PI * kx / NX);
            const double siny = sin(M_PI * ky / NY);
            const double denom = 4.0 * (sinx * sinx + siny * siny);
            const size_t k = (size_t)i * NYh + j;
            if (denom > 1e-15) { s->phi_hat[k].re = s->rho_hat[k].re / denom; s->phi_hat[k].im = s->rho_hat[k].im / denom; }
            else { s->phi_hat[k].re = 0.0; s->phi_hat[k].im = 0.0; }
        }
    offt_c2r_2d(s->fft, s->phi_hat, s->fft_out);
    const double norm = 1.0 / (int)N;
    for (size_t c = 0; c < N; ++c) s->phi[c] = s->fft_out[c] * norm;
}

/* poisson.cpp:90-142 (GS) and :216-279 (SOR): red-black sweeps over the interior, phi = 0 rim */
static void poisson_rb(po_state* s, int sor)
{
    const int NX = s->NX, NY = s->NY;
    double* phi = s->phi;
    const double omega = s->omega;
    for (size_t iter = 0; iter < 5000; ++iter) {                       /* maxIter, poisson.cpp:13 */
        double maxErr = 0.0;
        for (int colour = 0; colour < 2; ++colour) {
            #pragma omp parallel for reduction(max : maxErr) schedule(static)
            for (int j = 1; j < NY - 1; ++j)
                for (int i = 1; i < NX - 1; ++i) {
                    if (((i + j) & 1) != colour) continue;
                    const size_t c = ic(i, j, NX);
                    const double old = phi[c];
                    const double nb = phi[ic(i + 1, j, NX)] + phi[ic(i - 1, j, NX)] + phi[ic(i, j + 1, NX)] + phi[ic(i, j - 1, NX)];
                    const double gs = 0.25 * (nb + s->rho_q[c]);
                    const double nw = sor ? (1.0 - omega) * old + omega * gs : gs;
                    const double err = fabs(nw - old);
                    phi[c] = nw;
                    if (err > maxErr) maxErr = err;
                }
        }
        if (maxErr < 1e-8) break;                                      /* tol, poisson.cpp:14 */
    }
}

/* poisson.cpp:429-483: 4-colour 9-point Gauss-Seidel, colour = 2*(i&1) + (j&1) */
static void poisson_9pt(po_state* s)
{
    const int NX = s->NX, NY = s->NY;
    double* phi = s->phi;
    for (size_t iter = 0; iter < 5000; ++iter) {
        double maxErr = 0.0;
        for (int sweep = 0; sweep < 4; ++sweep) {
            #pragma omp parallel for reduction(max : maxErr) schedule(static)
            for (int j = 1; j < NY - 1; ++j)
                for (int i = 1; i < NX - 1; ++i) {
                    if ((2 * (i & 1) + (j & 1)) != sweep) continue;
                    const size_t c = ic(i, j, NX);
                    const double so = phi[ic(i + 1, j, NX)] + phi[ic(i - 1, j, NX)] + phi[ic(i, j + 1, NX)] + phi[ic(i, j - 1, NX)];
                    const double sd = phi[ic(i + 1, j + 1, NX)] + phi[ic(i - 1, j + 1, NX)] + phi[ic(i + 1, j - 1, NX)] + phi[ic(i - 1, j - 1, NX)];
                    const double nw = (4.0 * so + sd + 6.0 * s->rho_q[c]) / 20.0;
                    const double err = fabs(nw - phi[c]);
                    phi[c] = nw;
                    if (err > maxErr) maxErr = err;
                }
        }
        if (maxErr < 1e-8) break;
    }
}

/* poisson.cpp:589-607 */
static void efield_periodic(po_state* s)
{
    const int NX = s->NX, NY = s->NY;
    for (int j = 0; j < NY; ++j)
        for (int i = 0; i < NX; ++i) {
            const int im1 = (i + NX - 1) % NX, ip1 = (i + 1) % NX, jm1 = (j + NY - 1) % NY, jp1 = (j + 1) % NY;
            s->Ex[ic(i, j, NX)] = -0.5 * (s->phi[ic(ip1, j, NX)] - s->phi[ic(im1, j, NX)]);
            s->Ey[ic(i, j, NX)] = -0.5 * (s->phi[ic(i, jp1, NX)] - s->phi[ic(i, jm1, NX)]);
        }
}

/* poisson.cpp:551-585: interior central differences, then Neumann copy onto the rim
 * (rows first, then columns -- the order fixes the corner values) */
static void efield_walls(po_state* s)
{
    const int NX = s->NX, NY = s->NY;
    for (int j = 1; j < NY - 1; ++j)
        for (int i = 1; i < NX - 1; ++i) {
            s->Ex[ic(i, j, NX)] = -0.5 * (s->phi[ic(i + 1, j, NX)] - s->phi[ic(i - 1, j, NX)]);
            s->Ey[ic(i, j, NX)] = -0.5 * (s->phi[ic(i, j + 1, NX)] - s->phi[ic(i, j - 1, NX)]);
        }
    for (int i = 0; i < NX; ++i) {
        s->Ex[ic(i, 0, NX)] = s->Ex[ic(i, 1, NX)];           s->Ey[ic(i, 0, NX)] = s->Ey[ic(i, 1, NX)];
        s->Ex[ic(i, NY - 1, NX)] = s->Ex[ic(i, NY - 2, NX)]; s->Ey[ic(i, NY - 1, NX)] = s->Ey[ic(i, NY - 2, NX)];
    }
    for (int j = 0; j < NY; ++j) {
        s->Ex[ic(0, j, NX)] = s->Ex[ic(1, j, NX)];           s->Ey[ic(0, j, NX)] = s->Ey[ic(1, j, NX)];
        s->Ex[ic(NX - 1, j, NX)] = s->Ex[ic(NX - 2, j, NX)]; s->Ey[ic(NX - 1, j, NX)] = s->Ey[ic(NX - 2, j, NX)];
    }
}

/* poisson.cpp:25-82 */
void po_solve_poisson(po_state* s)
{
    const size_t N = (size_t)s->NX * s->NY;
    if (!s->poisson_called) {                                          /* call_once, :34-41 */
        s->poisson_called = 1;
        memset(s->phi, 0, sizeof(double) * N);
        if (s->poisson_type == PO_POISSON_NONE) {
            memset(s->Ex, 0, sizeof(double) * N);
            memset(s->Ey, 0, sizeof(double) * N);
        }
    }
    if (s->poisson_type == PO_POISSON_NONE) return;                    /* :43 */
    if (s->bc_type == PO_BC_PERIODIC) {                                /* :46-64 */
        switch (s->poisson_type) {
        case PO_POISSON_GS: poisson_rb(s, 0); break;
        case PO_POISSON_SOR: poisson_rb(s, 1); break;
        case PO_POISSON_NPS: poisson_9pt(s); break;
        case PO_POISSON_FFT: poisson_fft(s); break;
        default: return;
        }
        efield_periodic(s);
    } else {                                                           /* :65-80 */
        switch (s->poisson_type) {
        case PO_POISSON_GS: poisson_rb(s, 0); break;
        case PO_POISSON_SOR: poisson_rb(s, 1); break;
        case PO_POISSON_NPS: poisson_9pt(s); break;
        default: return;                                               /* FFT + walls: field untouched */
        }
        efield_walls(s);
    }
}

/* loop body of LBmethod::Run_simulation, plasma.cpp:476-513 */
void po_step(po_state* s, int nsteps)
{
    for (int t = 0; t < nsteps; ++t) {
        po_update_macro(s);
        po_compute_equilibrium(s);
        po_thermal_collisions(s);
        po_collisions(s);
        po_stream(s);
        po_solve_poisson(s);
    }
}
