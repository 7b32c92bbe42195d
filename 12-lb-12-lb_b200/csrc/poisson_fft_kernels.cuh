// poisson_fft_kernels.cuh -- what the three spectral passes share: the plan-shape dispatch and small helpers.
// Each pass lives in its own translation unit (poisson_rows_fwd.cu, poisson_cols.cu, poisson_rows_inv.cu) so that
// the ~30 kernel instantiations per pass compile in parallel.
#pragma once
#include "poisson_fft.h"
#include "fft.cuh"

#include <type_traits>

namespace plbm {

// FFT_CAP = the largest CTA the instantiation may be launched with (sets the register budget:
// 512 threads -> 128 registers, 768 -> 85 with spills; only sequences longer than 8192 need the latter).

// (A, B) of a row pair are neighbours in T[k][row]: one 256-bit access when the pair is 32-byte aligned
__device__ __forceinline__ void store_pair(cpx* p, cpx a, cpx b)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(a.re), "d"(a.im), "d"(b.re), "d"(b.im) : "memory");
}
__device__ __forceinline__ void load_pair(const cpx* p, cpx& a, cpx& b)
{
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.re), "=d"(a.im), "=d"(b.re), "=d"(b.im) : "l"(p));
}

inline size_t fft_smem_bytes(int n) { return sizeof(cpx) * (size_t)fft_smem_elems(n); }

// kernels are specialised on the plan's shape (tail pass kind, odd passes present) and on the CTA cap
template <class F>
inline cudaError_t with_shape(const FftPlan& P, F&& f)
{
#ifdef PLBM_FFT_ONE_SHAPE   // development builds: only the shape of power-of-two lengths 2^(4k+1) <= 8192 (e.g. 8192, 512), seconds to compile
    return f(std::integral_constant<int, 512>{}, std::integral_constant<int, FFT_TAIL_2>{}, std::integral_constant<int, FFT_ODD_NONE>{});
#else
    auto odd = [&](auto TAIL) {
        if (P.threads > 512) {
            if (P.odd != FFT_ODD_NONE) return f(std::integral_constant<int, 768>{}, TAIL, std::integral_constant<int, FFT_ODD_GENERIC>{});
            return f(std::integral_constant<int, 768>{}, TAIL, std::integral_constant<int, FFT_ODD_NONE>{});
        }
        switch (P.odd) {
        case FFT_ODD_3: return f(std::integral_constant<int, 512>{}, TAIL, std::integral_constant<int, FFT_ODD_3>{});
        case FFT_ODD_5: return f(std::integral_constant<int, 512>{}, TAIL, std::integral_constant<int, FFT_ODD_5>{});
        case FFT_ODD_35: return f(std::integral_constant<int, 512>{}, TAIL, std::integral_constant<int, FFT_ODD_35>{});
        case FFT_ODD_GENERIC: return f(std::integral_constant<int, 512>{}, TAIL, std::integral_constant<int, FFT_ODD_GENERIC>{});
        default: return f(std::integral_constant<int, 512>{}, TAIL, std::integral_constant<int, FFT_ODD_NONE>{});
        }
    };
    switch (P.tail) {
    case FFT_TAIL_42: return odd(std::integral_constant<int, FFT_TAIL_42>{});
    case FFT_TAIL_4: return odd(std::integral_constant<int, FFT_TAIL_4>{});
    case FFT_TAIL_2: return odd(std::integral_constant<int, FFT_TAIL_2>{});
    default: return odd(std::integral_constant<int, FFT_TAIL_NONE>{});
    }
#endif
}

template <class K>
inline cudaError_t allow_smem(K kernel)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fft_smem_bytes(FFT_MAX_N));
}

// per-pass halves of poisson_fft_configure
cudaError_t configure_poisson_rows_fwd(const PoissonFftDev& p);
cudaError_t configure_poisson_cols(const PoissonFftDev& p);
cudaError_t configure_poisson_rows_inv(const PoissonFftDev& p);

} // namespace plbm
