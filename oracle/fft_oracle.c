/*
 * oracle/fft_oracle.c -- TEST INFRASTRUCTURE ONLY.  See fft_oracle.h for the specification.
 * Stand-in for the FFTW3 calls of /root/reference/src/poisson.cpp:384,412,621-622.
 * Build with -ffp-contract=off: every rounding below is part of the specification.
 */
#include "fft_oracle.h"

#include <quadmath.h>
#include <stdlib.h>
#include <string.h>

int offt_factorize(int n, int* radix)
{
    int k = 0;
    while (n % 4 == 0) { radix[k++] = 4; n /= 4; }
    while (n % 2 == 0) { radix[k++] = 2; n /= 2; }
    for (int p = 3; p * p <= n; p += 2)
        while (n % p == 0) { radix[k++] = p; n /= p; }
    if (n > 1) radix[k++] = n;
    return k;
}

offt_plan1d* offt_plan1d_create(int n)
{
    offt_plan1d* p = (offt_plan1d*)calloc(1, sizeof(*p));
    p->n = n;
    p->nfac = offt_factorize(n, p->radix);
    p->tw = (offt_cpx*)malloc(sizeof(offt_cpx) * (size_t)n);
    const __float128 two_pi = 2.0Q * M_PIq;
    for (int t = 0; t < n; ++t) {
        const __float128 ang = two_pi * (__float128)t / (__float128)n;
        p->tw[t].re = (double)cosq(ang);
        p->tw[t].im = (double)(-sinq(ang));
    }
    return p;
}

void offt_plan1d_destroy(offt_plan1d* p)
{
    if (!p) return;
    free(p->tw);
    free(p);
}

static inline offt_cpx cadd(offt_cpx a, offt_cpx b) { offt_cpx r = { a.re + b.re, a.im + b.im }; return r; }
static inline offt_cpx csub(offt_cpx a, offt_cpx b) { offt_cpx r = { a.re - b.re, a.im - b.im }; return r; }
static inline offt_cpx cmul(offt_cpx a, offt_cpx w)
{
    offt_cpx r = { a.re * w.re - a.im * w.im, a.re * w.im + a.im * w.re };
    return r;
}
/* forward twiddle W[idx], conjugated when sign > 0 */
static inline offt_cpx twid(const offt_plan1d* p, long idx, int sign)
{
    offt_cpx w = p->tw[idx % p->n];
    if (sign > 0) w.im = -w.im;
    return w;
}

static void stage_radix2(const offt_plan1d* P, int sign, int n, int s, const offt_cpx* x, offt_cpx* y)
{
    const int m = n / 2;
    for (int p = 0; p < m; ++p) {
        const offt_cpx w1 = twid(P, (long)p * s, sign);
        for (int q = 0; q < s; ++q) {
            const offt_cpx a0 = x[q + s * (p + 0)];
            const offt_cpx a1 = x[q + s * (p + m)];
            const offt_cpx b0 = cadd(a0, a1);
            const offt_cpx b1 = csub(a0, a1);
            y[q + s * (2 * p + 0)] = b0;
            y[q + s * (2 * p + 1)] = (p == 0) ? b1 : cmul(b1, w1);
        }
    }
}

static void stage_radix4(const offt_plan1d* P, int sign, int n, int s, const offt_cpx* x, offt_cpx* y)
{
    const int m = n / 4;
    for (int p = 0; p < m; ++p) {
        const offt_cpx w1 = twid(P, (long)p * s, sign);
        const offt_cpx w2 = twid(P, 2L * p * s, sign);
        const offt_cpx w3 = twid(P, 3L * p * s, sign);
        for (int q = 0; q < s; ++q) {
            const offt_cpx a0 = x[q + s * (p + 0 * m)];
            const offt_cpx a1 = x[q + s * (p + 1 * m)];
            const offt_cpx a2 = x[q + s * (p + 2 * m)];
            const offt_cpx a3 = x[q + s * (p + 3 * m)];
            const offt_cpx t0 = cadd(a0, a2);
            const offt_cpx t1 = csub(a0, a2);
            const offt_cpx t2 = cadd(a1, a3);
            const offt_cpx t3 = csub(a1, a3);
            const offt_cpx b0 = cadd(t0, t2);
            const offt_cpx b2 = csub(t0, t2);
            offt_cpx b1, b3;
            if (sign < 0) { /* w4 = -i */
                b1.re = t1.re + t3.im; b1.im = t1.im - t3.re;
                b3.re = t1.re - t3.im; b3.im = t1.im + t3.re;
            } else {        /* w4 = +i */
                b1.re = t1.re - t3.im; b1.im = t1.im + t3.re;
                b3.re = t1.re + t3.im; b3.im = t1.im - t3.re;
            }
            offt_cpx* o = y + q + s * (4 * p);
            if (p == 0) {
                o[0] = b0; o[s] = b1; o[2 * s] = b2; o[3 * s] = b3;
            } else {
                o[0] = b0; o[s] = cmul(b1, w1); o[2 * s] = cmul(b2, w2); o[3 * s] = cmul(b3, w3);
            }
        }
    }
}

/* any odd prime radix: direct r-point DFT, terms accumulated in k order */
static void stage_generic(const offt_plan1d* P, int sign, int r, int n, int s, const offt_cpx* x, offt_cpx* y,
                          offt_cpx* a)
{
    const int m = n / r;
    const long step = P->n / r; /* w_r^e = W[e * N/r] */
    for (int p = 0; p < m; ++p) {
        for (int q = 0; q < s; ++q) {
            for (int k = 0; k < r; ++k) a[k] = x[q + s * (p + m * k)];
            for (int j = 0; j < r; ++j) {
                offt_cpx acc = a[0];
                for (int k = 1; k < r; ++k) {
                    if (j == 0) acc = cadd(acc, a[k]);
                    else        acc = cadd(acc, cmul(a[k], twid(P, (long)((j * (long)k) % r) * step, sign)));
                }
                if (j != 0 && p != 0) acc = cmul(acc, twid(P, (long)j * p * s, sign));
                y[q + s * (r * p + j)] = acc;
            }
        }
    }
}

offt_cpx* offt_exec1d(const offt_plan1d* P, int sign, offt_cpx* x, offt_cpx* work)
{
    int n = P->n, s = 1;
    offt_cpx* src = x;
    offt_cpx* dst = work;
    offt_cpx* scratch = NULL;
    for (int f = 0; f < P->nfac; ++f) {
        const int r = P->radix[f];
        if (r == 4) stage_radix4(P, sign, n, s, src, dst);
        else if (r == 2) stage_radix2(P, sign, n, s, src, dst);
        else {
            if (!scratch) scratch = (offt_cpx*)malloc(sizeof(offt_cpx) * (size_t)P->n);
            stage_generic(P, sign, r, n, s, src, dst, scratch);
        }
        n /= r; s *= r;
        offt_cpx* t = src; src = dst; dst = t;
    }
    free(scratch);
    return src;
}

offt_plan2d* offt_plan2d_create(int n0, int n1)
{
    offt_plan2d* p = (offt_plan2d*)calloc(1, sizeof(*p));
    p->n0 = n0; p->n1 = n1; p->nh = n1 / 2 + 1;
    p->row = offt_plan1d_create(n1);
    p->col = offt_plan1d_create(n0);
    return p;
}

void offt_plan2d_destroy(offt_plan2d* p)
{
    if (!p) return;
    offt_plan1d_destroy(p->row);
    offt_plan1d_destroy(p->col);
    free(p);
}

/* Row pass of the r2c transform on `nrows` consecutive rows (pairs (0,1),(2,3),... of the given
 * block): out[r][k], k = 0..n1/2.  Exposed separately so that the slab-decomposed solve can be
 * restated on the CPU (tests/slab_double.py). */
void offt_rows_fwd(const offt_plan2d* P, const double* in, int nrows, offt_cpx* out)
{
    const int n1 = P->n1, nh = P->nh;
    const int npairs = (nrows + 1) / 2;
    #pragma omp parallel
    {
        offt_cpx* z = (offt_cpx*)malloc(sizeof(offt_cpx) * (size_t)n1);
        offt_cpx* w = (offt_cpx*)malloc(sizeof(offt_cpx) * (size_t)n1);
        #pragma omp for schedule(static)
        for (int pr = 0; pr < npairs; ++pr) {
            const int ra = 2 * pr, rb = 2 * pr + 1;
            const int paired = rb < nrows;
            for (int j = 0; j < n1; ++j) {
                z[j].re = in[(size_t)ra * n1 + j];
                z[j].im = paired ? in[(size_t)rb * n1 + j] : 0.0;
            }
            const offt_cpx* Z = offt_exec1d(P->row, -1, z, w);
            if (!paired) {
                for (int k = 0; k < nh; ++k) out[(size_t)ra * nh + k] = Z[k];
            } else {
                for (int k = 0; k < nh; ++k) {
                    const int km = (n1 - k) % n1;
                    const double a = Z[k].re, b = Z[k].im, c = Z[km].re, d = Z[km].im;
                    out[(size_t)ra * nh + k].re = 0.5 * (a + c);
                    out[(size_t)ra * nh + k].im = 0.5 * (b - d);
                    out[(size_t)rb * nh + k].re = 0.5 * (b + d);
                    out[(size_t)rb * nh + k].im = 0.5 * (c - a);
                }
            }
        }
        free(z); free(w);
    }
}

/* One column transform of length n0, in place (sign -1 forward, +1 backward). */
void offt_col(const offt_plan2d* P, int sign, offt_cpx* col)
{
    const int n0 = P->n0;
    offt_cpx* w = (offt_cpx*)malloc(sizeof(offt_cpx) * (size_t)n0);
    offt_cpx* Z = offt_exec1d(P->col, sign, col, w);
    if (Z != col) memcpy(col, Z, sizeof(offt_cpx) * (size_t)n0);
    free(w);
}

/* Row pass of the c2r transform on `nrows` consecutive rows: H[r][k] -> out[r][j]. */
void offt_rows_inv(const offt_plan2d* P, const offt_cpx* H, int nrows, double* out)
{
    const int n1 = P->n1, nh = P->nh;
    const int npairs = (nrows + 1) / 2;
    #pragma omp parallel
    {
        offt_cpx* z = (offt_cpx*)malloc(sizeof(offt_cpx) * (size_t)n1);
        offt_cpx* w = (offt_cpx*)malloc(sizeof(offt_cpx) * (size_t)n1);
        #pragma omp for schedule(static)
        for (int pr = 0; pr < npairs; ++pr) {
            const int ra = 2 * pr, rb = 2 * pr + 1;
            const int paired = rb < nrows;
            const offt_cpx* Ha = H + (size_t)ra * nh;
            const offt_cpx* Hb = H + (size_t)(paired ? rb : ra) * nh;
            for (int k = 0; k < nh; ++k) {
                const int self_conj = (k == 0) || (2 * k == n1); /* DC / Nyquist: imaginary part ignored */
                double ar = Ha[k].re, ai = self_conj ? 0.0 : Ha[k].im;
                double br = paired ? Hb[k].re : 0.0;
                double bi = (paired && !self_conj) ? Hb[k].im : 0.0;
                z[k].re = ar - bi;
                z[k].im = ai + br;
                if (!self_conj) {
                    z[n1 - k].re = ar + bi;
                    z[n1 - k].im = br - ai;
                }
            }
            const offt_cpx* Z = offt_exec1d(P->row, +1, z, w);
            for (int j = 0; j < n1; ++j) {
                out[(size_t)ra * n1 + j] = Z[j].re;
                if (paired) out[(size_t)rb * n1 + j] = Z[j].im;
            }
        }
        free(z); free(w);
    }
}

void offt_r2c_2d(const offt_plan2d* P, const double* in, offt_cpx* out)
{
    const int n0 = P->n0, nh = P->nh;
    offt_rows_fwd(P, in, n0, out);
    #pragma omp parallel
    {
        offt_cpx* z = (offt_cpx*)malloc(sizeof(offt_cpx) * (size_t)n0);
        #pragma omp for schedule(static)
        for (int k = 0; k < nh; ++k) {
            for (int r = 0; r < n0; ++r) z[r] = out[(size_t)r * nh + k];
            offt_col(P, -1, z);
            for (int r = 0; r < n0; ++r) out[(size_t)r * nh + k] = z[r];
        }
        free(z);
    }
}

void offt_c2r_2d(const offt_plan2d* P, const offt_cpx* in, double* out)
{
    const int n0 = P->n0, nh = P->nh;
    offt_cpx* H = (offt_cpx*)malloc(sizeof(offt_cpx) * (size_t)n0 * nh);
    #pragma omp parallel
    {
        offt_cpx* z = (offt_cpx*)malloc(sizeof(offt_cpx) * (size_t)n0);
        #pragma omp for schedule(static)
        for (int k = 0; k < nh; ++k) {
            for (int r = 0; r < n0; ++r) z[r] = in[(size_t)r * nh + k];
            offt_col(P, +1, z);
            for (int r = 0; r < n0; ++r) H[(size_t)r * nh + k] = z[r];
        }
        free(z);
    }
    offt_rows_inv(P, H, n0, out);
    free(H);
}
