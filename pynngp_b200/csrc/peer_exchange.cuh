// peer_exchange.cuh -- the allreduce of the 3 statistics fused into the tail of the likelihood kernel.
//
// The path's only exchange step is the sum of {sum log F, sum r^2/F, n_bad} over the shards (24 bytes per
// parameter vector): pure latency.  Instead of a second launch (NCCL: ~20 us against a 55 us kernel at
// n = 1e6 on 8 GPUs) the last block of every rank stores its three numbers straight into every peer's
// exchange buffer over NVLink (P2P stores), waits for the lines of all ranks in its own buffer and sums
// them in rank order -- every rank obtains the bitwise identical total.
//
// Wire format (the low-latency line of NCCL's LL protocol, applied to doubles): one value = one 16-byte line
//     {lo32(value), stamp, hi32(value), stamp}
// written by a single 16-byte store.  Each 8-byte half carries the generation stamp beside its payload, and
// an aligned 8-byte store is never torn, so a reader that sees both stamps equal to the generation it waits
// for holds the whole value: no separate flag, no __threadfence_system between data and flag, no second
// round trip.  Buffers alternate with the generation's parity: a rank can be at most one generation ahead
// of the slowest one, because completing generation g needs every rank's lines for g.
// The same line format carries the result to the HOST (mapped pinned memory polled by the CPU).
#pragma once
#include "nngp_common.cuh"

__device__ __forceinline__ void ll_store(uint4 *dst, double v, unsigned int stamp)
{
    const unsigned int lo = (unsigned int)__double2loint(v), hi = (unsigned int)__double2hiint(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(lo), "r"(stamp), "r"(hi), "r"(stamp)
                 : "memory");
}
__device__ __forceinline__ bool ll_try_load(const uint4 *src, unsigned int stamp, double *v)
{
    unsigned int lo, s0, hi, s1;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(s0), "=r"(hi), "=r"(s1) : "l"(src) : "memory");
    *v = __hiloint2double(int(hi), int(lo));
    return s0 == stamp && s1 == stamp;
}

__host__ __device__ inline size_t peer_line_index(int par, int rank, int K_cap, int k, int comp)
{
    return ((size_t(par) * NNGP_MAX_PEERS + rank) * size_t(K_cap) + size_t(k)) * 3 + comp;
}

// Called by all threads of the (single) block that holds the rank's totals for parameter vector k; needs
// blockDim.x >= 3 * px.world.  v0..v2 are read from thread 0.  Returns the all-rank sums to thread 0 through
// out3 (NaN after a ~30 s timeout: a peer never arrived).
__device__ __forceinline__ void peer_allreduce3(const PeerExchange &px, int k, double v0, double v1, double v2,
                                                double *sh /* >= 3 + 2 * 24 doubles of shared memory */, double out3[3])
{
    const int par = int(px.gen & 1u);
    const int t = threadIdx.x;
    if (t == 0) { sh[0] = v0; sh[1] = v1; sh[2] = v2; }
    __syncthreads();
    if (t < 3 * px.world) {
        // thread (peer, comp): store this rank's component into the peer's buffer ...
        const int peer = t / 3, comp = t - 3 * peer;
        ll_store(px.lines[peer] + peer_line_index(par, px.rank, px.K_cap, k, comp), sh[comp], px.gen);
        // ... and wait for the peer's component in this rank's own buffer
        const uint4 *mine = px.lines[px.rank] + peer_line_index(par, peer, px.K_cap, k, comp);
        double v = 0.0;
        bool ok = ll_try_load(mine, px.gen, &v);
        if (!ok) {
            const long long t0 = clock64();
            while (!(ok = ll_try_load(mine, px.gen, &v)))
                if (clock64() - t0 > 60000000000ll) break;  // ~30 s: ranks may be launched seconds apart
        }
        sh[3 + t] = v;
        sh[3 + 24 + t] = ok ? 1.0 : 0.0;
    }
    __syncthreads();
    if (t == 0) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        bool ok = true;
        for (int r = 0; r < px.world; ++r) {  // rank order: every rank adds the same numbers in the same order
            ok &= sh[3 + 24 + 3 * r] != 0.0 && sh[3 + 24 + 3 * r + 1] != 0.0 && sh[3 + 24 + 3 * r + 2] != 0.0;
            s0 += sh[3 + 3 * r]; s1 += sh[3 + 3 * r + 1]; s2 += sh[3 + 3 * r + 2];
        }
        const double bad = __longlong_as_double(0x7ff8000000000000ll);
        out3[0] = ok ? s0 : bad; out3[1] = ok ? s1 : bad; out3[2] = ok ? s2 : bad;
    }
}

// thread 0 of the block that holds the final numbers of parameter vector k: to device memory, or as stamped
// lines to the mapped host buffer the CPU is polling
__device__ __forceinline__ void publish_result(const EvalArgs &a, int k, const double tot[3])
{
    if (a.out) {
        a.out[size_t(k) * 3 + 0] = tot[0];
        a.out[size_t(k) * 3 + 1] = tot[1];
        a.out[size_t(k) * 3 + 2] = tot[2];
    }
    if (a.hout) {
        ll_store(a.hout + size_t(k) * 3 + 0, tot[0], a.seq);
        ll_store(a.hout + size_t(k) * 3 + 1, tot[1], a.seq);
        ll_store(a.hout + size_t(k) * 3 + 2, tot[2], a.seq);
    }
}
