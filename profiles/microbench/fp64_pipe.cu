// fp64_pipe.cu -- FP64 pipe latency/throughput on sm_100a: cycles per DFMA for W warps per SM
// sub-partition and C independent dependency chains per thread.  nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void chains(double* out, int iters, long long* cycles)
{
    double a[C];
    #pragma unroll
    for (int c = 0; c < C; ++c) a[c] = 1.0 + threadIdx.x * 1e-9 + c;
    const double m = 1.0000001, b = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        #pragma unroll
        for (int k = 0; k < 8; ++k) {
            #pragma unroll
            for (int c = 0; c < C; ++c) a[c] = __fma_rn(a[c], m, b);
        }
    }
    long long t1 = clock64();
    double s = 0;
    #pragma unroll
    for (int c = 0; c < C; ++c) s += a[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int C>
void run(int warps_per_smsp)
{
    double* out; long long* cyc; long long h;
    const int threads = warps_per_smsp * 4 * 32;
    cudaMalloc(&out, sizeof(double) * threads * 148);
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 2000;
    chains<C><<<148, threads>>>(out, iters, cyc);
    chains<C><<<148, threads>>>(out, iters, cyc);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per_warp_instr = (double)h / (iters * 8.0 * C);
    printf("warps/SMSP %d chains %d: %.2f cycles per DFMA per warp, SMSP issue interval %.2f cycles/DFMA\n",
           warps_per_smsp, C, per_warp_instr, per_warp_instr / warps_per_smsp);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int w : {1, 2, 3, 4, 6, 8}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
