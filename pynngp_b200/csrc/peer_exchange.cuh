// peer_exchange.cuh -- the allreduce of the 3 statistics fused into the tail of the likelihood kernel.
//
// The path's only exchange step is the sum of {sum log F, sum r^2/F, n_bad} over the shards (24 bytes per
// parameter vector): pure latency.  Instead of a second launch (NCCL: ~20 us against a 55 us kernel at
// n = 1e6 on 8 GPUs) the last block of every rank stores its three numbers straight into every peer's
// exchange buffer over NVLink (P2P stores to CUDA-IPC mapped memory), raises a generation flag there, waits
// for the flags of all ranks in its own buffer and sums the slots in rank order -- every rank obtains the
// bitwise identical total.  Buffers alternate with the generation's parity: a rank can be at most one
// generation ahead of the slowest one, because completing generation g needs every rank's flag for g.
#pragma once
#include "nngp_common.cuh"

// Called by all threads of the (single) block that holds the rank's totals for parameter vector k; needs
// blockDim.x >= px.world.  v0..v2 are read from thread 0.  Returns the all-rank sums to thread 0 through
// out3 (NaN after a ~30 s timeout: a peer never arrived).
__device__ __forceinline__ void peer_allreduce3(const PeerExchange &px, int k, double v0, double v1, double v2,
                                                double *sh /* >= 3 + 8 doubles of shared memory */, double out3[3])
{
    const int par = int(px.gen & 1ull);
    const int t = threadIdx.x;
    if (t == 0) { sh[0] = v0; sh[1] = v1; sh[2] = v2; }
    __syncthreads();
    if (t < px.world) {
        // thread t serves peer t: store this rank's slot there, then the flag
        volatile double *ps = px.slots[t] + (size_t(par) * NNGP_MAX_PEERS + px.rank) * size_t(px.K_cap) * 3 + size_t(k) * 3;
        ps[0] = sh[0]; ps[1] = sh[1]; ps[2] = sh[2];
        __threadfence_system();
        volatile unsigned long long *pf = px.flags[t] + (size_t(par) * NNGP_MAX_PEERS + px.rank) * size_t(px.K_cap) + k;
        *pf = px.gen;
        // ... and waits for rank t's flag in this rank's own buffer
        volatile unsigned long long *mf = px.flags[px.rank] + (size_t(par) * NNGP_MAX_PEERS + t) * size_t(px.K_cap) + k;
        const long long t0 = clock64();
        bool ok = true;
        while (*mf != px.gen) {
            if (clock64() - t0 > 60000000000ll) { ok = false; break; }  // ~30 s: ranks may be launched seconds apart
            __nanosleep(64);
        }
        __threadfence_system();
        sh[3 + t] = ok ? 1.0 : 0.0;
    }
    __syncthreads();
    if (t == 0) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        bool ok = true;
        for (int r = 0; r < px.world; ++r) {
            ok &= sh[3 + r] != 0.0;
            volatile const double *ms = px.slots[px.rank] + (size_t(par) * NNGP_MAX_PEERS + r) * size_t(px.K_cap) * 3 + size_t(k) * 3;
            s0 += ms[0]; s1 += ms[1]; s2 += ms[2];
        }
        const double bad = __longlong_as_double(0x7ff8000000000000ll);
        out3[0] = ok ? s0 : bad; out3[1] = ok ? s1 : bad; out3[2] = ok ? s2 : bad;
    }
}
