// Developer microbenchmark: FP64 FMA dependent-issue latency and pipe throughput on sm_100a as a function of
// (warps per SM sub-partition, independent chains per thread).   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(int iters, double seed, double* out, long long* cyc)
{
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = seed + threadIdx.x + i;
    const double m = 0.999999, c = 1e-6;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    if (s == -12345.0) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP>
void run(int warps_per_sm, int sms)
{
    double* out; long long* cyc; long long h = 0;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    const int iters = 20000;
    k<ILP><<<sms, warps_per_sm * 32>>>(iters, 1.0, out, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<ILP><<<sms, warps_per_sm * 32>>>(iters, 1.0, out, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double fmas_per_warp = (double)iters * 8 * ILP;
    const double cyc_per_fma_per_warp = (double)h / fmas_per_warp;
    const double tf = 2.0 * fmas_per_warp * 32 * warps_per_sm * sms / (ms * 1e-3) / 1e12;
    // per SMSP: warps_per_sm/4 warps, each issuing one DFMA every cyc_per_fma cycles
    printf("warps/SM %2d  ILP %d : %.2f cycles per DFMA per warp, %.2f warp-DFMA/cycle/SMSP, %.2f TFLOP/s\n", warps_per_sm, ILP,
           cyc_per_fma_per_warp, (warps_per_sm / 4.0) / cyc_per_fma_per_warp, tf);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    for (int w : {4, 8, 12, 16, 32}) {
        run<1>(w, sms); run<2>(w, sms); run<3>(w, sms); run<4>(w, sms); run<8>(w, sms);
    }
    return 0;
}
