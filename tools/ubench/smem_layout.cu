// Shared-memory wavefront cost of the fused kernel's per-location buffers under candidate paddings (sm_100a).
// A warp carries W = 32 / G locations; location g's buffer starts at
//     off(g) = g * base + (g & 1) * a + ((g >> 1) & 1) * b + (g >> 2) * c      (bytes)
// and the kernel's accesses are: column publish (8-byte store, lane (g, q) -> row of block s), pivot read (8-byte
// broadcast load inside a location), column read (16-byte broadcast load), staged-point write (16-byte store per row),
// staged-point read (16-byte broadcast load).  Prints cycles per warp-instruction at SM level (8 warps) = wavefronts.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench/smem_layout.cu -o tools/ubench/smem_layout
//   tools/ubench/smem_layout G R elem_bytes base a b c [base a b c ...]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int THREADS = 256;
constexpr int UNROLL = 16;
enum Acc { ST8_ROWS, LD8_BCAST, LD16_BCAST, ST16_ROWS, NACC };
const char *kAcc[NACC] = {"8B store rows", "8B bcast load", "16B bcast load", "16B store rows"};

template <int ACC>
__global__ void __launch_bounds__(THREADS) k(int G, int R, int elem, int base, int a, int b, int c, int iters, long long *cyc, uint32_t *sink)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 16384; i += THREADS) reinterpret_cast<uint32_t *>(sm)[i] = i;
    __syncthreads();
    const int g = lane / G, q = lane % G, P = G * R;
    const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(sm)) + warp * 8192 +
                           g * base + (g & 1) * a + ((g >> 1) & 1) * b + (g >> 2) * c;
    uint32_t off[UNROLL];
    for (int u = 0; u < UNROLL; ++u) {
        const int s = u % R;
        const int row = (s & 1) ? s * G + (G - 1 - q) : s * G + q;  // folded row assignment
        if (ACC == ST8_ROWS || ACC == ST16_ROWS) off[u] = sbase + row * elem;
        else off[u] = sbase + ((u * 2) % P) * elem;
    }
    uint32_t iacc = lane;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (ACC == ST8_ROWS) asm volatile("st.volatile.shared.v2.u32 [%0], {%1,%2};" ::"r"(off[u]), "r"(iacc), "r"(iacc) : "memory");
            else if (ACC == ST16_ROWS) asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%2,%1,%2};" ::"r"(off[u]), "r"(iacc), "r"(iacc) : "memory");
            else if (ACC == LD8_BCAST) { uint2 v; asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(off[u])); iacc ^= v.x ^ v.y; }
            else { uint4 v; asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(off[u])); iacc ^= v.x ^ v.w; }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (iacc == 0x12345678u) sink[0] = iacc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ACC>
double run(int sms, int G, int R, int elem, int base, int a, int b, int c)
{
    long long *cyc; uint32_t *sink;
    cudaMalloc(&cyc, 8 * sms); cudaMalloc(&sink, 4);
    const int iters = 1000;
    cudaFuncSetAttribute(k<ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<ACC><<<sms, THREADS, 65536>>>(G, R, elem, base, a, b, c, 10, cyc, sink);
    k<ACC><<<sms, THREADS, 65536>>>(G, R, elem, base, a, b, c, iters, cyc, sink);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, 8 * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += double(h[i]);
    cudaFree(cyc); cudaFree(sink);
    return mean / sms / (double(iters) * UNROLL * (THREADS / 32));
}

int main(int argc, char **argv)
{
    if (argc < 8) { printf("usage: smem_layout G R elem_bytes base a b c [base a b c ...]\n"); return 1; }
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const int G = atoi(argv[1]), R = atoi(argv[2]), elem = atoi(argv[3]);
    for (int i = 4; i + 3 < argc; i += 4) {
        const int base = atoi(argv[i]), a = atoi(argv[i + 1]), b = atoi(argv[i + 2]), c = atoi(argv[i + 3]);
        printf("G=%d R=%d elem=%d off(g)=g*%d+(g&1)*%d+(g>>1&1)*%d+(g>>2)*%d :", G, R, elem, base, a, b, c);
        if (elem == 8) printf("  st8 %.2f  ld8b %.2f  ld16b %.2f\n", run<ST8_ROWS>(sms, G, R, elem, base, a, b, c), run<LD8_BCAST>(sms, G, R, elem, base, a, b, c),
                              run<LD16_BCAST>(sms, G, R, elem, base, a, b, c));
        else printf("  st16 %.2f  ld16b %.2f\n", run<ST16_ROWS>(sms, G, R, elem, base, a, b, c), run<LD16_BCAST>(sms, G, R, elem, base, a, b, c));
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
