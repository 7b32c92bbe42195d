// fft.cuh -- register-tiled Stockham FFT core used by the Poisson kernels (poisson_fft.cu).
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
// One CTA transforms one length-n sequence; every thread owns 16 elements.  Two consecutive
// radix-4 stages of the schedule (or 4 then 2) are executed as ONE pass on a thread's 16 registers:
// the thread that holds x[b + (n/16)*kappa], kappa = 0..15, can form the four first-stage
// butterflies p = p2 + m2*k2 and then the four second-stage butterflies that consume exactly their
// outputs, so the arithmetic per output -- and therefore every bit -- is that of the stage-by-stage
// definition while the sequence crosses shared memory once per PASS instead of once per stage.
// The first pass reads through a caller-supplied functor (global memory, coalesced: consecutive
// threads read consecutive elements) and the last pass writes through one, so a 2048-point
// transform (4,4 | 4,4 | 4,2) makes two trips through shared memory.  Odd prime stages come last
// in the schedule and run as the definition says, one output per (thread, slot).
// Shared memory is padded by one element per 16 (slot()) so that the first pass's stride-16
// stores spread over all banks.
#pragma once
#include "exact_math.cuh"

namespace plbm {

struct __align__(16) cpx {
    double re, im;
};

constexpr int FFT_MAX_PASSES = 12;   // n <= 12288: at most 9 odd stages, or 7 stages of 4/2 in 4 passes
constexpr int FFT_EPT = 16;         // elements per thread
constexpr int FFT_MAX_N = 12288;    // 768 threads * 16; 204 KB of padded shared memory

enum FftPassKind { FFT_PASS_44 = 0, FFT_PASS_42 = 1, FFT_PASS_4 = 2, FFT_PASS_2 = 3, FFT_PASS_ODD = 4 };

struct FftPass {
    unsigned short r;        // FFT_PASS_ODD: the prime
    unsigned short s;        // stride = product of the radices of all earlier stages (a power of two for the 4/2 passes)
    unsigned short m;        // sub-length after the pass: nsub / R
    unsigned char kind;
    unsigned char log2s;
};

struct FftPlan {
    int n;
    int npass;
    int n44;                        // leading FFT_PASS_44 passes
    int tail;                       // FftTail: the 4/2 pass after them
    int odd;                        // FftOdd
    int n3;                         // FFT_ODD_35: number of radix-3 stages (they precede the radix-5 stages)
    int threads;                    // CTA size the passes need: ceil(n/16), or ceil(n/15) with radix-3/5 register passes
    FftPass pass[FFT_MAX_PASSES];
    const cpx* tw;                  // device: n forward twiddles exp(-2 pi i t / n)
};

__host__ __device__ __forceinline__ int fft_slot(int i) { return i + (i >> 4); }
__host__ __device__ __forceinline__ int fft_smem_elems(int n) { return fft_slot(n > 0 ? n - 1 : 0) + 1; }
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return { __dadd_rn(a.re, b.re), __dadd_rn(a.im, b.im) }; }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return { __dsub_rn(a.re, b.re), __dsub_rn(a.im, b.im) }; }
__device__ __forceinline__ cpx cmul(cpx a, cpx w)
{
    return { __dsub_rn(__dmul_rn(a.re, w.re), __dmul_rn(a.im, w.im)),
             __dadd_rn(__dmul_rn(a.re, w.im), __dmul_rn(a.im, w.re)) };
}
template <int SIGN>
__device__ __forceinline__ cpx twiddle(const cpx* __restrict__ tw, int idx)
{
    const double2 t = __ldg(reinterpret_cast<const double2*>(tw) + idx);   // one 16-byte load
    cpx w;
    w.re = t.x;
    w.im = (SIGN > 0) ? -t.y : t.y;
    return w;
}

// radix-4 / radix-2 butterflies in place, before the stage twiddle
template <int SIGN>
__device__ __forceinline__ void bfly4(cpx& a0, cpx& a1, cpx& a2, cpx& a3)
{
    const cpx t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    if (SIGN < 0) {
        a1 = { __dadd_rn(t1.re, t3.im), __dsub_rn(t1.im, t3.re) };
        a3 = { __dsub_rn(t1.re, t3.im), __dadd_rn(t1.im, t3.re) };
    } else {
        a1 = { __dsub_rn(t1.re, t3.im), __dadd_rn(t1.im, t3.re) };
        a3 = { __dadd_rn(t1.re, t3.im), __dsub_rn(t1.im, t3.re) };
    }
}
__device__ __forceinline__ void bfly2(cpx& a0, cpx& a1)
{
    const cpx d = csub(a0, a1);
    a0 = cadd(a0, a1);
    a1 = d;
}

// Source of a pass: the caller's functor for the first pass, the (padded) shared buffer afterwards; likewise the
// destination: shared memory until the last pass.  One type per transform keeps ONE copy of every pass body in the
// kernel (the choice is a CTA-uniform branch per access), which matters more than the branch: the bodies are long
// straight-line code and the instruction cache is a real resource here.
template <class First>
struct FftSource {
    const First& f; const cpx* buf; bool first;
    __device__ __forceinline__ bool is_smem() const { return !first || First::is_smem; }
    __device__ __forceinline__ cpx load(int i) const { return first ? f.load(i) : buf[fft_slot(i)]; }
};
template <class Last>
struct FftSink {
    const Last& l; cpx* buf; bool last;
    __device__ __forceinline__ bool is_smem() const { return !last || Last::is_smem; }
    __device__ __forceinline__ void store(int i, cpx v) const
    {
        if (last) l.store(i, v);
        else buf[fft_slot(i)] = v;
    }
};
// plain shared-memory endpoints for callers (buf must hold fft_smem_elems(n) elements)
struct FftSmem {
    cpx* buf;
    static constexpr bool is_smem = true;
    __device__ __forceinline__ cpx load(int i) const { return buf[fft_slot(i)]; }
    __device__ __forceinline__ void store(int i, cpx v) const { buf[fft_slot(i)] = v; }
};

// One pass = stage of radix R1 followed (R2 > 1) by the stage of radix R2, on registers.
//   inputs   x[b + nb*(k2 + R2*k)]           b = q + s*p2 the butterfly index, nb = n/(R1*R2)
//   stage 1  p = p2 + m*k2:  y_j = (sum_k x_k w^(jk)) * W[j*p*s]      (twiddle skipped when p == 0, as defined)
//   stage 2  stride s*R1:    z_j2 = (sum_k2 y_k2 w^(j2 k2)) * W[j2*p2*s*R1]
//   outputs  z[q + s*(j + R1*j2 + R1*R2*p2)]
template <int SIGN, int R1, int R2, class In, class Out>
__device__ __forceinline__ void fft_pass_pow2(int n, const FftPass ps, const cpx* __restrict__ tw, const In& in, const Out& out)
{
    constexpr int R = R1 * R2, BPT = FFT_EPT / R;
    const int nb = n / R, nt = blockDim.x, s = ps.s;
    cpx v[BPT][R1][R2];
    #pragma unroll
    for (int e = 0; e < BPT; ++e) {
        const int b = threadIdx.x + e * nt;
        if (b < nb) {
            #pragma unroll
            for (int k = 0; k < R1; ++k)
                #pragma unroll
                for (int k2 = 0; k2 < R2; ++k2) v[e][k][k2] = in.load(b + nb * (k2 + R2 * k));
            const int ps_ = (b >> ps.log2s) << ps.log2s;          // p2 * s
            #pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) {
                if constexpr (R1 == 4) bfly4<SIGN>(v[e][0][k2], v[e][1][k2], v[e][2][k2], v[e][3][k2]);
                else bfly2(v[e][0][k2], v[e][1][k2]);
                const int t1 = ps_ + nb * k2;                     // p * s  (m * s == nb)
                if (k2 != 0 || t1 != 0) {                          // p != 0
                    #pragma unroll
                    for (int j = 1; j < R1; ++j) v[e][j][k2] = cmul(v[e][j][k2], twiddle<SIGN>(tw, j * t1));
                }
            }
            if constexpr (R2 > 1) {
                #pragma unroll
                for (int j = 0; j < R1; ++j) {
                    if constexpr (R2 == 4) bfly4<SIGN>(v[e][j][0], v[e][j][1], v[e][j][2], v[e][j][3]);
                    else bfly2(v[e][j][0], v[e][j][1]);
                }
                if (ps_ != 0) {
                    #pragma unroll
                    for (int j2 = 1; j2 < R2; ++j2) {
                        const cpx w = twiddle<SIGN>(tw, j2 * ps_ * R1);
                        #pragma unroll
                        for (int j = 0; j < R1; ++j) v[e][j][j2] = cmul(v[e][j][j2], w);
                    }
                }
            }
        }
    }
    if (in.is_smem() && out.is_smem()) __syncthreads();           // one buffer: every read before any write
    #pragma unroll
    for (int e = 0; e < BPT; ++e) {
        const int b = threadIdx.x + e * nt;
        if (b < nb) {
            const int p2 = b >> ps.log2s, q = b - (p2 << ps.log2s);
            const int o = q + s * (R * p2);
            #pragma unroll
            for (int j2 = 0; j2 < R2; ++j2)
                #pragma unroll
                for (int j = 0; j < R1; ++j) out.store(o + s * (j + R1 * j2), v[e][j][j2]);
        }
    }
    if (out.is_smem()) __syncthreads();
}

// odd prime radix: every (thread, slot) forms one whole output, terms accumulated in k order
template <int SIGN, class In, class Out>
__device__ __forceinline__ void fft_pass_odd(int n, const FftPass ps, const cpx* __restrict__ tw, const In& in, const Out& out)
{
    const int r = ps.r, s = ps.s, m = ps.m, nt = blockDim.x;
    const int step = n / r;                                       // w_r^e = W[e * n / r]
    cpx acc_[FFT_EPT];
    #pragma unroll
    for (int e = 0; e < FFT_EPT; ++e) {
        const int o = threadIdx.x + e * nt;
        if (o < n) {
            const int pj = o / s, q = o - pj * s;                 // pj = r*p + j
            const int p = pj / r, j = pj - p * r;
            cpx acc = in.load(q + s * p);
            for (int k = 1; k < r; ++k) {
                const cpx ak = in.load(q + s * (p + m * k));
                if (j == 0) acc = cadd(acc, ak);
                else acc = cadd(acc, cmul(ak, twiddle<SIGN>(tw, ((j * k) % r) * step)));
            }
            if (j != 0 && p != 0) acc = cmul(acc, twiddle<SIGN>(tw, j * p * s));
            acc_[e] = acc;
        }
    }
    if (in.is_smem() && out.is_smem()) __syncthreads();
    #pragma unroll
    for (int e = 0; e < FFT_EPT; ++e) {
        const int o = threadIdx.x + e * nt;
        if (o < n) out.store(o, acc_[e]);
    }
    if (out.is_smem()) __syncthreads();
}

// radix 3 or 5 on registers: a thread forms whole butterflies (all R outputs from R inputs), 16/R of them,
// with exactly the generic pass's arithmetic per output (terms in k order, w_R^e taken from the table).
template <int SIGN, int R, class In, class Out>
__device__ __forceinline__ void fft_pass_small_odd(int n, const FftPass ps, const cpx* __restrict__ tw, const In& in, const Out& out)
{
    constexpr int BPT = FFT_EPT / R;
    const int nb = n / R, nt = blockDim.x, s = ps.s;
    cpx w[R];                                                     // w[e] = w_R^e = W[e * n / R]
    #pragma unroll
    for (int e = 1; e < R; ++e) w[e] = twiddle<SIGN>(tw, e * nb);
    cpx v[BPT][R];
    #pragma unroll
    for (int e = 0; e < BPT; ++e) {
        const int b = threadIdx.x + e * nt;                       // b = q + s*p
        if (b < nb) {
            cpx a[R];
            #pragma unroll
            for (int k = 0; k < R; ++k) a[k] = in.load(b + nb * k);   // s*m == nb
            const int ps_ = (b / s) * s;                          // p * s
            #pragma unroll
            for (int j = 0; j < R; ++j) {
                cpx acc = a[0];
                #pragma unroll
                for (int k = 1; k < R; ++k) {
                    if (j == 0) acc = cadd(acc, a[k]);
                    else acc = cadd(acc, cmul(a[k], w[(j * k) % R]));
                }
                if (j != 0 && ps_ != 0) acc = cmul(acc, twiddle<SIGN>(tw, j * ps_));
                v[e][j] = acc;
            }
        }
    }
    if (in.is_smem() && out.is_smem()) __syncthreads();
    #pragma unroll
    for (int e = 0; e < BPT; ++e) {
        const int b = threadIdx.x + e * nt;
        if (b < nb) {
            const int p = b / s, q = b - p * s;
            #pragma unroll
            for (int j = 0; j < R; ++j) out.store(q + s * (R * p + j), v[e][j]);
        }
    }
    if (out.is_smem()) __syncthreads();
}

// Shape of a plan that the kernels are specialised on: the schedule is always [4,4]* then at most one of
// (4,2) / (4) / (2), then odd primes.  Keeping the pass bodies out of a runtime switch lets ptxas allocate each
// body's 16 complex registers independently (a switch over the bodies made it spill).
enum FftTail { FFT_TAIL_NONE = 0, FFT_TAIL_42 = 1, FFT_TAIL_4 = 2, FFT_TAIL_2 = 3 };
// odd part of the schedule: none; 3 / 5 = every odd stage has that radix, 15 = radix 3 and 5 mixed (register passes);
// generic otherwise
enum FftOdd { FFT_ODD_NONE = 0, FFT_ODD_GENERIC = 1, FFT_ODD_3 = 3, FFT_ODD_5 = 5, FFT_ODD_35 = 15 };   // 15: stages of radix 3 and 5 mixed

// Transform of length P.n from `in` to `out` (element index -> value functors with load/store and a static
// `is_smem`), through the padded shared buffer `buf`.  All threads of the CTA must call it.  If `in` reads shared
// memory, its contents must be visible (barrier) before the call; if `out` writes shared memory, the result is
// visible to every thread on return.  TAIL / ODD must describe P (P.tail, P.odd).
template <int SIGN, int TAIL, int ODD, class In, class Out>
__device__ __forceinline__ void fft_run(const FftPlan& P, cpx* buf, const In& in, const Out& out)
{
    if (P.npass == 0) {                                           // n == 1
        if (threadIdx.x < P.n) out.store(threadIdx.x, in.load(threadIdx.x));
        if (Out::is_smem) __syncthreads();
        return;
    }
    const int last = P.npass - 1;
    int i = 0;
    #pragma unroll 1
    for (; i < P.n44; ++i) {
        const FftSource<In> src{ in, buf, i == 0 };
        const FftSink<Out> dst{ out, buf, i == last };
        fft_pass_pow2<SIGN, 4, 4>(P.n, P.pass[i], P.tw, src, dst);
    }
    if constexpr (TAIL != FFT_TAIL_NONE) {
        const FftSource<In> src{ in, buf, i == 0 };
        const FftSink<Out> dst{ out, buf, i == last };
        if constexpr (TAIL == FFT_TAIL_42) fft_pass_pow2<SIGN, 4, 2>(P.n, P.pass[i], P.tw, src, dst);
        else if constexpr (TAIL == FFT_TAIL_4) fft_pass_pow2<SIGN, 4, 1>(P.n, P.pass[i], P.tw, src, dst);
        else fft_pass_pow2<SIGN, 2, 1>(P.n, P.pass[i], P.tw, src, dst);
        ++i;
    }
    if constexpr (ODD == FFT_ODD_35) {
        // primes ascend in the schedule: all radix-3 stages, then all radix-5 stages -- two loops, one body each
        const int end3 = i + P.n3;
        #pragma unroll 1
        for (; i < end3; ++i) {
            const FftSource<In> src{ in, buf, i == 0 };
            const FftSink<Out> dst{ out, buf, i == last };
            fft_pass_small_odd<SIGN, 3>(P.n, P.pass[i], P.tw, src, dst);
        }
        #pragma unroll 1
        for (; i < P.npass; ++i) {
            const FftSource<In> src{ in, buf, i == 0 };
            const FftSink<Out> dst{ out, buf, i == last };
            fft_pass_small_odd<SIGN, 5>(P.n, P.pass[i], P.tw, src, dst);
        }
    } else if constexpr (ODD != FFT_ODD_NONE) {
        #pragma unroll 1
        for (; i < P.npass; ++i) {
            const FftSource<In> src{ in, buf, i == 0 };
            const FftSink<Out> dst{ out, buf, i == last };
            if constexpr (ODD == 3 || ODD == 5) fft_pass_small_odd<SIGN, ODD>(P.n, P.pass[i], P.tw, src, dst);
            else fft_pass_odd<SIGN>(P.n, P.pass[i], P.tw, src, dst);
        }
    }
}

} // namespace plbm
