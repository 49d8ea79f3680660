// Microbenchmarks that size the fused kernel's TLP/ILP needs on sm_100a:
//  (1) dependent DFMA latency (one warp, one chain)   (2) FP64 throughput vs resident warps x ILP
//  (3) MUFU.RSQ64H, SHFL, LDS dependent latencies.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_chain(double *out, int iters, double a, double b, long long *cyc)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == -1.2345) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void shfl_chain(double *out, int iters, long long *cyc)
{
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 16; ++u) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 3, 4) + 1.0;
    long long t1 = clock64();
    if (x == -1.2345) out[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void rsq_chain(double *out, int iters, long long *cyc)
{
    double x = 1.0 + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            double y;
            asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
            x = y + 1.5;
        }
    long long t1 = clock64();
    if (x == -1.2345) out[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP>
void run_tp(int warps_per_sm, int sms)
{
    double *out; long long *cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    int threads = warps_per_sm * 32;  // one block per SM
    int iters = 4096;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    dfma_chain<ILP><<<sms, threads>>>(out, 64, 0.999, 1e-7, cyc);
    cudaEventRecord(a);
    dfma_chain<ILP><<<sms, threads>>>(out, iters, 0.999, 1e-7, cyc);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double instr = double(sms) * threads * iters * 16.0 * ILP;
    printf("warps/SM=%2d ILP=%d: %.3f Tinstr/s  cycles/warp-instr(per warp)=%.2f\n", warps_per_sm, ILP,
           instr / (ms * 1e-3) / 1e12, double(c) / (iters * 16.0 * ILP));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s SMs=%d\n", p.name, sms);
    double *out; long long *cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    long long c;
    dfma_chain<1><<<1, 32>>>(out, 1024, 0.999, 1e-7, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent latency: %.2f cycles\n", double(c) / (1024 * 16.0));
    shfl_chain<<<1, 32>>>(out, 1024, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("SHFL.f64(2x)+DADD dependent latency: %.2f cycles\n", double(c) / (1024 * 16.0));
    rsq_chain<<<1, 32>>>(out, 1024, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("MUFU.RSQ64H+DADD dependent latency: %.2f cycles\n", double(c) / (1024 * 16.0));
    for (int w : {4, 8, 12, 16, 24, 32}) { run_tp<1>(w, sms); }
    for (int w : {4, 8, 12, 16}) { run_tp<2>(w, sms); }
    for (int w : {4, 8, 12, 16}) { run_tp<4>(w, sms); }
    return 0;
}
