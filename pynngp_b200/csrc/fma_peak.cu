// fma_peak.cu -- measures the device's FP64 / FP32 FMA issue peak (thread-instructions per second),
// the denominator of the fused kernel's roofline (MEASURED_PEAKS.json carries no vector-pipe figure).
// Eight independent register-resident FMA chains per thread; every SM filled with resident warps.
#include "nngp_common.cuh"

namespace {

template <typename T>
__global__ void __launch_bounds__(256) fma_chain_kernel(T *sink, int iters, T a, T b)
{
    T x0 = T(threadIdx.x), x1 = x0 + T(1), x2 = x0 + T(2), x3 = x0 + T(3);
    T x4 = x0 + T(4), x5 = x0 + T(5), x6 = x0 + T(6), x7 = x0 + T(7);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
            x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
        }
    }
    const T s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == T(-12345.678)) sink[0] = s;  // never true; keeps the chains alive
}

template <typename T>
cudaError_t run(nngp_handle *h, int iters, double *instr_per_s)
{
    T *sink = nullptr;
    cudaError_t e = cudaMalloc(&sink, sizeof(T));
    if (e != cudaSuccess) return e;
    const int grid = h->num_sms * 8, block = 256;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    fma_chain_kernel<T><<<grid, block, 0, h->stream>>>(sink, iters / 8 + 1, T(0.999999), T(1e-7));  // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(t0, h->stream);
        fma_chain_kernel<T><<<grid, block, 0, h->stream>>>(sink, iters, T(0.999999), T(1e-7));
        cudaEventRecord(t1, h->stream);
        if ((e = cudaEventSynchronize(t1)) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        const double instr = double(grid) * block * double(iters) * 64.0;
        const double rate = instr / (ms * 1e-3);
        if (rate > best) best = rate;
        h->launches += 1;
    }
    *instr_per_s = best;
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(sink);
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace

cudaError_t launch_fma_peak(nngp_handle *h, int dtype, int iters, double *instr_per_s)
{
    return dtype == NNGP_F64 ? run<double>(h, iters, instr_per_s) : run<float>(h, iters, instr_per_s);
}
