/*
 * oracle/fft_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Deterministic CPU FFT that stands in for FFTW3, which the reference links
 * (-lfftw3, /root/reference/compile_and_run_plasma.sh:32; call sites
 * /root/reference/src/poisson.cpp:384,412,616-622) but which is NOT vendored in the
 * reference tree and is not installed in this image (third-party, version unpinned:
 * /root/reference/README.md:13 "libfftw3-dev").  "Parity unpinned" at that boundary: the
 * reference holds no golden vectors for it, so the transform is restated here from the
 * published definition FFTW documents (unnormalised DFT, row-major n0 x n1, half spectrum
 * n1/2+1 along the contiguous axis) and cross-checked against numpy.fft in
 * tests/test_oracle_fft.py.
 *
 * The algorithm is FIXED (it is the specification the CUDA FFT in
 * 12-lb-12-lb_b200/csrc/fft.cu mirrors bit for bit):
 *   - 1-D complex DFT, Stockham autosort, decimation in frequency, radix schedule
 *     4,4,..,2,(odd primes ascending); stage with sub-length n, stride s, m = n/r:
 *         a_k = x[q + s*(p + m*k)],  b_j = sum_k a_k w_r^(jk),  y[q + s*(r*p + j)] = b_j * W[(j*p*s) mod N]
 *   - twiddles W[t] = exp(-2*pi*i*t/N) rounded from __float128 (libquadmath cosq/sinq),
 *     conjugated for the backward transform; no FMA contraction anywhere (-ffp-contract=off).
 *   - r2c: two real rows are packed into one complex transform and split by Hermitian symmetry;
 *     an unpaired last row is transformed alone with zero imaginary part.
 */
#ifndef PLBM_FFT_ORACLE_H
#define PLBM_FFT_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } offt_cpx;

typedef struct offt_plan1d {
    int n;            /* transform length */
    int nfac;         /* number of stages */
    int radix[64];    /* radix of each stage, in execution order */
    offt_cpx* tw;     /* n forward twiddles exp(-2 pi i t / n) */
} offt_plan1d;

offt_plan1d* offt_plan1d_create(int n);
void offt_plan1d_destroy(offt_plan1d* p);
/* radix schedule only (shared rule with the CUDA implementation's host planner tests) */
int offt_factorize(int n, int* radix);
/* sign = -1 forward, +1 backward (unnormalised).  x and work have n entries; the result is
 * returned in whichever of the two buffers the last stage wrote: the returned pointer. */
offt_cpx* offt_exec1d(const offt_plan1d* p, int sign, offt_cpx* x, offt_cpx* work);

/* 2-D real <-> half-complex transforms with FFTW's r2c_2d / c2r_2d conventions:
 * in: n0 rows of n1 contiguous reals; out: n0 rows of (n1/2+1) complex. */
typedef struct offt_plan2d {
    int n0, n1, nh;
    offt_plan1d* row;   /* length n1 */
    offt_plan1d* col;   /* length n0 */
} offt_plan2d;

offt_plan2d* offt_plan2d_create(int n0, int n1);
void offt_plan2d_destroy(offt_plan2d* p);
void offt_r2c_2d(const offt_plan2d* p, const double* in, offt_cpx* out);
/* the separable pieces (row pass on a block of rows, one column transform in place) */
void offt_rows_fwd(const offt_plan2d* p, const double* in, int nrows, offt_cpx* out);
void offt_rows_inv(const offt_plan2d* p, const offt_cpx* H, int nrows, double* out);
void offt_col(const offt_plan2d* p, int sign, offt_cpx* col);
void offt_c2r_2d(const offt_plan2d* p, const offt_cpx* in, double* out);

#ifdef __cplusplus
}
#endif
#endif
