// fft.cuh -- shared-memory Stockham FFT core used by the Poisson kernels (poisson_fft.cu).
//
// The reference solves Poisson with FFTW r2c/c2r plans (/root/reference/src/poisson.cpp:384,412,
// 621-622).  FFTW is a third-party library that is absent from the reference tree, so there is
// no "reference rounding" to match; the transform is therefore DEFINED once (see
// oracle/fft_oracle.h for the specification: radix schedule 4..,2,odd primes; DIF Stockham stage
//     a_k = x[q + s*(p + m*k)],  y[q + s*(r*p + j)] = (sum_k a_k w_r^(jk)) * W[j*p*s]
// with a forward twiddle table rounded from binary128) and implemented identically on the CPU
// checker and here, with the same operation order and no FMA, so that potential and field are
// bit-identical to the checker's.
//
// One CTA transforms one length-n sequence held in a single shared-memory buffer: in every stage
// each thread first reads the inputs of all its butterflies into registers, the CTA synchronises,
// and only then are outputs written (the Stockham permutation is not in place).
#pragma once
#include "exact_math.cuh"

namespace plbm {

struct cpx {
    double re, im;
};

constexpr int FFT_MAX_STAGES = 32;
constexpr int FFT_MAX_EPT = 12;     // elements per thread and stage
constexpr int FFT_MAX_N = 12288;    // 1024 threads * 12

struct FftPlan {
    int n;
    int nstages;
    int radix[FFT_MAX_STAGES];
    const cpx* tw;                  // device: n forward twiddles exp(-2 pi i t / n)
};

__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return { __dadd_rn(a.re, b.re), __dadd_rn(a.im, b.im) }; }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return { __dsub_rn(a.re, b.re), __dsub_rn(a.im, b.im) }; }
__device__ __forceinline__ cpx cmul(cpx a, cpx w)
{
    return { __dsub_rn(__dmul_rn(a.re, w.re), __dmul_rn(a.im, w.im)),
             __dadd_rn(__dmul_rn(a.re, w.im), __dmul_rn(a.im, w.re)) };
}
template <int SIGN, bool TW_SHARED = false>
__device__ __forceinline__ cpx twiddle(const cpx* __restrict__ tw, int idx)
{
    const double2 t = TW_SHARED ? reinterpret_cast<const double2*>(tw)[idx]
                                : __ldg(reinterpret_cast<const double2*>(tw) + idx);   // one 16-byte load
    cpx w;
    w.re = t.x;
    w.im = (SIGN > 0) ? -t.y : t.y;
    return w;
}

// In-place (single buffer) transform of buf[0..n).  All threads of the CTA must call it; on return
// the result is in natural order and visible to every thread.  EPT = elements per thread and stage
// (a multiple of 4): the CTA must have at least ceil(n / EPT) threads.
template <int SIGN, int EPT = FFT_MAX_EPT, bool TW_SHARED = false>
__device__ __forceinline__ void fft_smem(cpx* buf, const FftPlan& P, const cpx* tw_table)
{
    const int n = P.n;
    const int tid = threadIdx.x, nt = blockDim.x;
    int nsub = n, s = 1;
    __syncthreads();
    for (int st = 0; st < P.nstages; ++st) {
        const int r = P.radix[st];
        const int m = nsub / r;
        cpx out[EPT];
        if (r == 4) {
            const int nb = n >> 2;
            #pragma unroll
            for (int e = 0; e < EPT / 4; ++e) {
                const int b = tid + e * nt;
                if (b < nb) {
                    const int p = b / s, q = b - p * s;
                    const cpx a0 = buf[q + s * (p)];
                    const cpx a1 = buf[q + s * (p + m)];
                    const cpx a2 = buf[q + s * (p + 2 * m)];
                    const cpx a3 = buf[q + s * (p + 3 * m)];
                    const cpx t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
                    cpx b0 = cadd(t0, t2), b2 = csub(t0, t2), b1, b3;
                    if (SIGN < 0) {
                        b1 = { __dadd_rn(t1.re, t3.im), __dsub_rn(t1.im, t3.re) };
                        b3 = { __dsub_rn(t1.re, t3.im), __dadd_rn(t1.im, t3.re) };
                    } else {
                        b1 = { __dsub_rn(t1.re, t3.im), __dadd_rn(t1.im, t3.re) };
                        b3 = { __dadd_rn(t1.re, t3.im), __dsub_rn(t1.im, t3.re) };
                    }
                    if (p != 0) {
                        b1 = cmul(b1, twiddle<SIGN, TW_SHARED>(tw_table, p * s));
                        b2 = cmul(b2, twiddle<SIGN, TW_SHARED>(tw_table, 2 * p * s));
                        b3 = cmul(b3, twiddle<SIGN, TW_SHARED>(tw_table, 3 * p * s));
                    }
                    out[4 * e + 0] = b0; out[4 * e + 1] = b1; out[4 * e + 2] = b2; out[4 * e + 3] = b3;
                }
            }
            __syncthreads();
            #pragma unroll
            for (int e = 0; e < EPT / 4; ++e) {
                const int b = tid + e * nt;
                if (b < nb) {
                    const int p = b / s, q = b - p * s;
                    cpx* o = buf + q + s * (4 * p);
                    o[0] = out[4 * e + 0]; o[s] = out[4 * e + 1]; o[2 * s] = out[4 * e + 2]; o[3 * s] = out[4 * e + 3];
                }
            }
        } else if (r == 2) {
            const int nb = n >> 1;
            #pragma unroll
            for (int e = 0; e < EPT / 2; ++e) {
                const int b = tid + e * nt;
                if (b < nb) {
                    const int p = b / s, q = b - p * s;
                    const cpx a0 = buf[q + s * (p)];
                    const cpx a1 = buf[q + s * (p + m)];
                    cpx b1 = csub(a0, a1);
                    if (p != 0) b1 = cmul(b1, twiddle<SIGN, TW_SHARED>(tw_table, p * s));
                    out[2 * e + 0] = cadd(a0, a1); out[2 * e + 1] = b1;
                }
            }
            __syncthreads();
            #pragma unroll
            for (int e = 0; e < EPT / 2; ++e) {
                const int b = tid + e * nt;
                if (b < nb) {
                    const int p = b / s, q = b - p * s;
                    cpx* o = buf + q + s * (2 * p);
                    o[0] = out[2 * e + 0]; o[s] = out[2 * e + 1];
                }
            }
        } else {
            // odd prime radix: every thread forms whole outputs, terms accumulated in k order
            const int step = n / r;   // w_r^e = W[e * n / r]
            #pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int o = tid + e * nt;
                if (o < n) {
                    const int q = o % s;
                    const int pj = o / s;          // = r*p + j
                    const int p = pj / r, j = pj - p * r;
                    cpx acc = buf[q + s * p];
                    for (int k = 1; k < r; ++k) {
                        const cpx ak = buf[q + s * (p + m * k)];
                        if (j == 0) acc = cadd(acc, ak);
                        else acc = cadd(acc, cmul(ak, twiddle<SIGN, TW_SHARED>(tw_table, ((j * k) % r) * step)));
                    }
                    if (j != 0 && p != 0) acc = cmul(acc, twiddle<SIGN, TW_SHARED>(tw_table, j * p * s));
                    out[e] = acc;
                }
            }
            __syncthreads();
            #pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int o = tid + e * nt;
                if (o < n) buf[o] = out[e];
            }
        }
        __syncthreads();
        nsub = m; s *= r;
    }
}

} // namespace plbm
