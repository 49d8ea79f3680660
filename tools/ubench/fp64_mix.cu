// Issue-slot cost of FP64: DFMA alone vs DFMA interleaved 1:1 and 1:2 with independent ALU/FMA-pipe work.
#include <cstdio>
#include <cuda_runtime.h>

template <int NI, int NF>
__global__ void k(double *out, int iters, const double *in)
{
    double x[8];
    int a[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = in[i] + threadIdx.x; a[i] = threadIdx.x + i; f[i] = threadIdx.x * 0.5f + i; }
    const double y = in[8], z = in[16];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                x[i] = fma(x[i], y, z);
#pragma unroll
                for (int t = 0; t < NI; ++t) a[i] = (a[i] ^ (a[i] >> 3)) + it;   // LOP3/SHF/IADD (ALU pipe)
#pragma unroll
                for (int t = 0; t < NF; ++t) f[i] = fmaf(f[i], 0.999f, 0.5f);    // FFMA (FMA pipe)
            }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + a[i] + f[i];
    if (s == -1.2345) out[0] = s;
}

template <int NI, int NF>
void run(const char *name, int sms, double *in)
{
    double *out; cudaMalloc(&out, 8);
    int iters = 2048, threads = 256, blocks = sms * 2;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<NI, NF><<<blocks, threads>>>(out, 32, in);
    cudaEventRecord(a);
    k<NI, NF><<<blocks, threads>>>(out, iters, in);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-34s %.3f T dfma/s   (%.3f ms)\n", name, double(blocks) * threads * iters * 64.0 / (ms * 1e-3) / 1e12, ms);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double h[24]; for (int i = 0; i < 24; ++i) h[i] = 0.999 + 1e-4 * i;
    double *in; cudaMalloc(&in, sizeof(h)); cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    int s = p.multiProcessorCount;
    run<0, 0>("dfma only", s, in);
    run<0, 1>("dfma + 1 ffma", s, in);
    run<0, 2>("dfma + 2 ffma", s, in);
    run<0, 3>("dfma + 3 ffma", s, in);
    run<1, 0>("dfma + 1 int-op group (~3 instr)", s, in);
    run<1, 1>("dfma + int group + 1 ffma", s, in);
    return 0;
}
