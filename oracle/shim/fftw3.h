/*
 * oracle/shim/fftw3.h -- TEST INFRASTRUCTURE ONLY.
 * Minimal stand-in for the FFTW3 API subset the reference calls
 * (/root/reference/include/poisson.hpp:9; /root/reference/src/poisson.cpp:384,412,616-634),
 * because libfftw3 is not installed in this image.  Semantics follow FFTW's documented
 * conventions (unnormalised, row-major n0 x n1, half spectrum along n1); the arithmetic is the
 * fixed algorithm of oracle/fft_oracle.c.
 */
#ifndef PLBM_ORACLE_FFTW3_SHIM_H
#define PLBM_ORACLE_FFTW3_SHIM_H

#include <stdlib.h>
#include "../fft_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];

typedef struct fftw_shim_plan_s {
    offt_plan2d* p2d;
    int kind; /* 0 = r2c, 1 = c2r */
    double* real;
    fftw_complex* cpx;
} *fftw_plan;

#define FFTW_ESTIMATE (1U << 6)

static inline void* fftw_malloc(size_t n) { return malloc(n); }
static inline void fftw_free(void* p) { free(p); }

static inline fftw_plan fftw_plan_dft_r2c_2d(int n0, int n1, double* in, fftw_complex* out, unsigned flags)
{
    (void)flags;
    fftw_plan p = (fftw_plan)malloc(sizeof(*p));
    p->p2d = offt_plan2d_create(n0, n1); p->kind = 0; p->real = in; p->cpx = out;
    return p;
}
static inline fftw_plan fftw_plan_dft_c2r_2d(int n0, int n1, fftw_complex* in, double* out, unsigned flags)
{
    (void)flags;
    fftw_plan p = (fftw_plan)malloc(sizeof(*p));
    p->p2d = offt_plan2d_create(n0, n1); p->kind = 1; p->real = out; p->cpx = in;
    return p;
}
static inline void fftw_execute(const fftw_plan p)
{
    if (p->kind == 0) offt_r2c_2d(p->p2d, p->real, (offt_cpx*)p->cpx);
    else              offt_c2r_2d(p->p2d, (const offt_cpx*)p->cpx, p->real);
}
static inline void fftw_destroy_plan(fftw_plan p)
{
    if (!p) return;
    offt_plan2d_destroy(p->p2d);
    free(p);
}

#ifdef __cplusplus
}
#endif
#endif
