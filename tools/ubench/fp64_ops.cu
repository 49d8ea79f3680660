// Does the FP64 pipe sustain its peak with three distinct 64-bit source operands per DFMA?
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double *out, int iters, const double *in)
{
    double x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = in[i] + threadIdx.x; y[i] = in[8 + i]; z[i] = in[16 + i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) x[i] = fma(x[i], y[0], z[0]);          // shared multiplier/addend
                if (MODE == 1) x[i] = fma(x[i], y[i], z[i]);          // three distinct registers
                if (MODE == 2) x[i] = fma(x[i], y[(i + u) & 7], z[(i + 3 * u) & 7]);  // rotating operands
                if (MODE == 3) x[i] = x[i] * y[i];                    // DMUL two operands
                if (MODE == 4) x[i] = x[i] + y[i];                    // DADD
            }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == -1.2345) out[0] = s;
}

template <int MODE>
void run(const char *name, int sms, double *in)
{
    double *out; cudaMalloc(&out, 8);
    int iters = 2048, threads = 256, blocks = sms * 2;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, threads>>>(out, 32, in);
    cudaEventRecord(a);
    k<MODE><<<blocks, threads>>>(out, iters, in);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-28s %.3f Tinstr/s\n", name, double(blocks) * threads * iters * 64.0 / (ms * 1e-3) / 1e12);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double h[24]; for (int i = 0; i < 24; ++i) h[i] = 0.999 + 1e-4 * i;
    double *in; cudaMalloc(&in, sizeof(h)); cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0>("dfma shared operands", p.multiProcessorCount, in);
    run<1>("dfma 3 distinct operands", p.multiProcessorCount, in);
    run<2>("dfma rotating operands", p.multiProcessorCount, in);
    run<3>("dmul", p.multiProcessorCount, in);
    run<4>("dadd", p.multiProcessorCount, in);
    return 0;
}
