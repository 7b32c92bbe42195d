// host_tables.h -- see host_tables.cpp
#pragma once
namespace plbm {
void host_twiddles(int n, double* re_im_pairs);
void host_sin2_rows(int NX, double* sx2);
void host_sin2_cols(int NY, double* sy2);
int host_factorize(int n, int* radix);   // radix schedule 4..,2,odd primes ascending
}
