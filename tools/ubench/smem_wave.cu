// Shared-memory throughput of the access patterns the fused kernel uses (sm_100a): cycles per warp-instruction at
// SM level, i.e. the number of 128-byte wavefronts the data stage spends on each.  One block of 8 warps per SM (the
// kernel's residency), every warp issues independent loads/stores back to back.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench/smem_wave.cu -o tools/ubench/smem_wave && tools/ubench/smem_wave
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int THREADS = 256;
constexpr int UNROLL = 16;

enum Pat {
    LDS128_GROUP4,   // 8 groups of 4 lanes, a group reads ONE 16-byte address, groups 144 B apart (elimination column)
    LDS128_GROUP16,  // 2 groups of 16 lanes (the m = 30 shape), groups 272 B apart
    LDS128_ALL,      // all 32 lanes read one 16-byte address
    LDS128_DISTINCT, // 32 distinct consecutive 16-byte addresses (record reads)
    LDS64_GROUP4,    // a group reads one 8-byte address, groups 144 B apart (pivot, w_k)
    LDS64_DISTINCT,  // 32 consecutive 8-byte addresses
    LDS64_RAND2048,  // random entry of a 2048-entry table of doubles (the exp table)
    LDS64_RAND256,
    LDS64_RAND32,
    LDS64_RAND16,
    LDS32_RAND2048,  // one 4-byte word of a random entry (a table split into hi / lo words costs two of these)
    LDS32_RAND32,
    STS64_COL,       // lane (g, q) stores 8 bytes at g*144 + (4s+q)*8 (column publish)
    SHFL32_W4,       // 32-bit shuffle, width 4
    MIX_LDS64_SHFL,  // one LDS.64 (32 distinct) + one SHFL.32 per step: 3.2 if they share the data stage, ~2 if not
    MIX_LDS64_LDS64, // two LDS.64 (32 distinct) per step (4.0)
    NPAT
};
const char *kNames[NPAT] = {"LDS.128 group-of-4 broadcast (stride 144 B)", "LDS.128 group-of-16 broadcast", "LDS.128 all-lanes broadcast",
                            "LDS.128 32 distinct (512 B)", "LDS.64 group-of-4 broadcast", "LDS.64 32 distinct (256 B)",
                            "LDS.64 random of 2048", "LDS.64 random of 256", "LDS.64 random of 32", "LDS.64 random of 16",
                            "LDS.32 random of 2048 (8 B stride)", "LDS.32 random of 32 (8 B stride)", "STS.64 column publish", "SHFL.32 width 4", "LDS.64 distinct + SHFL.32 (per pair)", "LDS.64 distinct x 2 (per pair)"};

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int PAT>
__global__ void __launch_bounds__(THREADS) k(double *out, int iters, long long *cyc)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 6144; i += THREADS) reinterpret_cast<double *>(sm)[i] = 1.0 + i;
    __syncthreads();
    unsigned char *wb = sm + warp * 4096;  // per-warp region for the structured patterns
    uint32_t off[UNROLL];
    uint32_t seed = threadIdx.x * 2654435761u + blockIdx.x;
    for (int u = 0; u < UNROLL; ++u) {
        const int g4 = lane >> 2, q4 = lane & 3, g16 = lane >> 4;
        switch (PAT) {
        case LDS128_GROUP4: off[u] = warp * 4096 + g4 * 144 + (u % 8) * 16; break;
        case LDS128_GROUP16: off[u] = warp * 4096 + g16 * 272 + (u % 16) * 16; break;
        case LDS128_ALL: off[u] = warp * 4096 + u * 16; break;
        case LDS128_DISTINCT: off[u] = warp * 4096 + lane * 16 + (u % 4) * 512; break;
        case LDS64_GROUP4: off[u] = warp * 4096 + g4 * 144 + (u % 16) * 8; break;
        case LDS64_DISTINCT: off[u] = warp * 4096 + lane * 8 + (u % 8) * 256; break;
        case LDS64_RAND2048: off[u] = (lcg(seed) % 2048) * 8; break;
        case LDS64_RAND256: off[u] = (lcg(seed) % 256) * 8; break;
        case LDS64_RAND32: off[u] = (lcg(seed) % 32) * 8; break;
        case LDS64_RAND16: off[u] = (lcg(seed) % 16) * 8; break;
        case LDS32_RAND2048: off[u] = (lcg(seed) % 2048) * 8; break;
        case LDS32_RAND32: off[u] = (lcg(seed) % 32) * 8; break;
        case MIX_LDS64_SHFL: case MIX_LDS64_LDS64: off[u] = warp * 4096 + lane * 8 + (u % 8) * 256; break;
        case STS64_COL: off[u] = warp * 4096 + g4 * 144 + ((u % 4) * 4 + q4) * 8; break;
        default: off[u] = 0;
        }
    }
    (void)wb;
    double acc0 = 0, acc1 = 0;
    uint32_t iacc = lane;
    const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(sm));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (PAT == LDS128_GROUP4 || PAT == LDS128_GROUP16 || PAT == LDS128_ALL || PAT == LDS128_DISTINCT) {
                uint4 v;
                asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sbase + off[u]));
                iacc ^= v.x ^ v.w;
            } else if (PAT == LDS32_RAND2048 || PAT == LDS32_RAND32) {
                uint32_t v;
                asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(sbase + off[u]));
                iacc ^= v;
            } else if (PAT == STS64_COL) {
                asm volatile("st.volatile.shared.v2.u32 [%0], {%1,%2};" :: "r"(sbase + off[u]), "r"(iacc), "r"(iacc) : "memory");
            } else if (PAT == MIX_LDS64_SHFL || PAT == MIX_LDS64_LDS64) {
                uint2 v;
                asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(sbase + off[u]));
                iacc ^= v.x ^ v.y;
                if (PAT == MIX_LDS64_SHFL) iacc ^= __shfl_sync(0xffffffffu, seed + u + it, (lane + u) & 3, 4);
                else {
                    asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(sbase + off[(u + 3) % UNROLL]));
                    iacc ^= v.x ^ v.y;
                }
            } else if (PAT == SHFL32_W4) {
                iacc ^= __shfl_sync(0xffffffffu, seed + u + it, (lane + u) & 3, 4);
            } else {
                uint2 v;
                asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(sbase + off[u]));
                iacc ^= v.x ^ v.y;
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (iacc == 0x12345678u) out[0] = acc0 + acc1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int PAT>
void run(int sms)
{
    double *out; long long *cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * sms);
    const int iters = 2000;
    cudaFuncSetAttribute(k<PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
    k<PAT><<<sms, THREADS, 49152>>>(out, 10, cyc);
    k<PAT><<<sms, THREADS, 49152>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, 8 * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += double(h[i]);
    mean /= sms;
    const double per = mean / (double(iters) * UNROLL * (THREADS / 32));
    printf("%-48s %6.2f cycles per warp-instruction (SM level, 8 warps)\n", kNames[PAT], per);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    run<LDS128_GROUP4>(sms); run<LDS128_GROUP16>(sms); run<LDS128_ALL>(sms); run<LDS128_DISTINCT>(sms);
    run<LDS64_GROUP4>(sms); run<LDS64_DISTINCT>(sms);
    run<LDS64_RAND2048>(sms); run<LDS64_RAND256>(sms); run<LDS64_RAND32>(sms); run<LDS64_RAND16>(sms);
    run<LDS32_RAND2048>(sms); run<LDS32_RAND32>(sms); run<STS64_COL>(sms); run<SHFL32_W4>(sms); run<MIX_LDS64_SHFL>(sms); run<MIX_LDS64_LDS64>(sms);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
