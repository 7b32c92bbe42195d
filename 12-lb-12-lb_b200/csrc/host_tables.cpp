// host_tables.cpp -- host-side tables of the spectral Poisson solve (compiled by g++, not nvcc,
// because the twiddles are rounded from binary128).
//
//   twiddles     W[t] = exp(-2 pi i t / n), each component the double nearest to the binary128
//                value of libquadmath's cosq/sinq -- the definition the CPU checker
//                (oracle/fft_oracle.c) uses, restated here independently.
//   sin2 tables  sin^2(pi*k/N) exactly as the reference forms them with libm's sin
//                (/root/reference/src/poisson.cpp:391-395): sinx = std::sin(M_PI * kx / NX),
//                kx = i for i <= NX/2 else i - NX; siny = std::sin(M_PI * ky / NY).
#include "host_tables.h"

#include <cmath>
#include <quadmath.h>

namespace plbm {

void host_twiddles(int n, double* re_im_pairs)
{
    const __float128 two_pi = 2.0Q * M_PIq;
    for (int t = 0; t < n; ++t) {
        const __float128 ang = two_pi * (__float128)t / (__float128)n;
        re_im_pairs[2 * t + 0] = (double)cosq(ang);
        re_im_pairs[2 * t + 1] = (double)(-sinq(ang));
    }
}

void host_sin2_rows(int NX, double* sx2)
{
    for (int i = 0; i < NX; ++i) {
        const int kx = (i <= NX / 2) ? i : i - NX;
        const double sinx = std::sin(M_PI * kx / NX);
        sx2[i] = sinx * sinx;
    }
}

void host_sin2_cols(int NY, double* sy2)
{
    const int NYh = NY / 2 + 1;
    for (int j = 0; j < NYh; ++j) {
        const double siny = std::sin(M_PI * j / NY);
        sy2[j] = siny * siny;
    }
}

int host_factorize(int n, int* radix)
{
    int k = 0;
    while (n % 4 == 0) { radix[k++] = 4; n /= 4; }
    while (n % 2 == 0) { radix[k++] = 2; n /= 2; }
    for (int p = 3; p * p <= n; p += 2)
        while (n % p == 0) { radix[k++] = p; n /= p; }
    if (n > 1) radix[k++] = n;
    return k;
}

} // namespace plbm
