// knn_ordered.cu -- stage 1 of the NNGP hot path: the ordered k-nearest-neighbour search (sm_100a).
//
// Takes over _make_s_neighbor_sets, pyNNGP/nngp.py:49-62: the reference rebuilds a scikit-learn
// KDTree on s[0:i] for every i (nngp.py:55) and queries k = min(m, i) (nngp.py:57-61).  Here it is an
// exact triangular brute force:
//   - one thread per query row i, NNGP_KNN_TILE consecutive rows per block;
//   - candidate rows j < i stream through shared memory as contiguous 32-byte records
//     {x, y, z, yval}, moved by 1-D bulk async copies (cp.async.bulk -> SASS UBLKCP, the TMA engine)
//     into a two-stage ring completed through mbarriers, so the copy of tile c+1 overlaps the scan
//     of tile c;
//   - every thread reads the same candidate (shared-memory broadcast) and evaluates the squared
//     distance exactly as scikit-learn does -- fp64, one dimension after the other, products and
//     sums rounded separately (no FMA): sklearn/metrics/_dist_metrics.pxd.tp:39-49 -- so the
//     selection is bit-exact;
//   - the m best (d2, j) of a row live in a per-thread column of shared memory, kept sorted by
//     insertion; candidates arrive in ascending j and insertion is strict, which makes the total
//     order (d2, j): ties go to the smaller index;
//   - query tiles are handed out heaviest first from an atomic counter (work grows with i).
// Bound: the FP64 pipe (5 arithmetic + 1 compare instruction per pair in 2-D); candidates are
// read from HBM/L2 once per 128 queries.
#include <math.h>
#include <stdint.h>

#include "nngp_common.cuh"

namespace nngp_knn {

constexpr int TQ = NNGP_KNN_TILE;  // queries (threads) per block
constexpr int TC = 512;            // candidates per stage
constexpr int NSTAGE = 2;

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    // bounded: a lost completion traps instead of hanging the device
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 24)) __trap();
}
// global -> shared 1-D bulk copy (TMA engine), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

struct TopM {
    double *keys;  // [m][TQ], this thread's column
    int32_t *ids;
    int m;
    int cnt;
    double worst;  // keys[m-1] once full, +inf before

    __device__ __forceinline__ void insert(double d2, int32_t j)
    {
        int pos = cnt < m ? cnt : m - 1;
        while (pos > 0 && d2 < keys[(pos - 1) * TQ]) {  // strict: equal d2 keeps the earlier j first
            keys[pos * TQ] = keys[(pos - 1) * TQ];
            ids[pos * TQ] = ids[(pos - 1) * TQ];
            --pos;
        }
        keys[pos * TQ] = d2;
        ids[pos * TQ] = j;
        if (cnt < m) ++cnt;
        if (cnt == m) worst = keys[(m - 1) * TQ];
    }
};

// squared distance, scikit-learn order of operations, no contraction
template <bool DIM3>
__device__ __forceinline__ double dist2_sk(double qx, double qy, double qz, const double2 *rec)
{
    const double2 a = rec[0];
    const double dx = qx - a.x, dy = qy - a.y;
    double d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    if (DIM3) {
        const double dz = qz - rec[1].x;
        d = __dadd_rn(d, __dmul_rn(dz, dz));
    }
    return d;
}

// ORDERED: candidates are the predecessors j < i (the NNGP neighbour sets).  !ORDERED: every row j,
// the query itself included -- the plain k-NN behind the reference's `ws` warm start (nngp.py:45-47).
template <bool DIM3, bool ORDERED>
__global__ void __launch_bounds__(TQ) knn_ordered_kernel(const double4 *__restrict__ pts, int64_t n,
                                                         int m, int ntiles, int tile_offset,
                                                         int tile_stride, int first_tile, int64_t out_lo, int64_t cand_cap,
                                                         unsigned int *tile_counter,
                                                         int32_t *__restrict__ out)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double4 *cand = reinterpret_cast<double4 *>(smem_raw);  // NSTAGE x TC records
    double *keys_all = reinterpret_cast<double *>(smem_raw + size_t(NSTAGE) * TC * sizeof(double4));
    int32_t *ids_all = reinterpret_cast<int32_t *>(keys_all + size_t(m) * TQ);
    __shared__ __align__(8) uint64_t full_bar[NSTAGE];
    __shared__ int s_tile;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0u;  // bit s = parity of stage s's next completion (block-uniform)

    for (;;) {
        if (tid == 0) s_tile = (int)atomicAdd(tile_counter, 1u);
        __syncthreads();
        const int64_t rank = int64_t(tile_offset) + int64_t(s_tile) * tile_stride;
        if (rank >= ntiles - first_tile) break;  // tiles below first_tile are not wanted
        const int64_t tile = ntiles - 1 - rank;  // heaviest (largest i) first
        const int64_t q0 = tile * TQ;
        const int64_t i = q0 + tid;
        const bool live = i < n;
        const int64_t last_q = (q0 + TQ < n ? q0 + TQ : n) - 1;  // largest live query of the tile
        // candidates 0 .. last_q-1, never beyond cand_cap (or all rows)
        const int64_t ncand = ORDERED ? (last_q < cand_cap ? last_q : cand_cap) : n;
        const int nct = int((ncand + TC - 1) / TC);

        double qx = 0.0, qy = 0.0, qz = 0.0;
        if (live) {
            const double4 p = pts[i];
            qx = p.x; qy = p.y; qz = p.z;
        }
        TopM top;
        top.keys = keys_all + tid;
        top.ids = ids_all + tid;
        top.m = m;
        top.cnt = live ? 0 : m;
        top.worst = live ? INFINITY : -INFINITY;  // dead threads never insert

        auto issue = [&](int c) {
            const int s = c % NSTAGE;
            const int64_t c0 = int64_t(c) * TC;
            const uint32_t cnt = uint32_t(ncand - c0 < TC ? ncand - c0 : TC);
            const uint32_t bytes = cnt * uint32_t(sizeof(double4));
            mbar_expect_tx(&full_bar[s], bytes);
            bulk_g2s(cand + size_t(s) * TC, pts + c0, bytes, &full_bar[s]);
        };
        if (tid == 0) {
            if (nct > 0) issue(0);
            if (nct > 1) issue(1);
        }

        for (int c = 0; c < nct; ++c) {
            const int s = c % NSTAGE;
            const int64_t c0 = int64_t(c) * TC;
            const int jn = int(ncand - c0 < TC ? ncand - c0 : TC);
            mbar_wait(&full_bar[s], (phase >> s) & 1u);
            phase ^= 1u << s;
            const double2 *rec = reinterpret_cast<const double2 *>(cand + size_t(s) * TC);

            if ((!ORDERED || c0 + jn <= q0) && (jn & 3) == 0) {
                // every candidate precedes every query of the tile: no j < i test
                for (int jj = 0; jj < jn; jj += 4) {
                    const double d0 = dist2_sk<DIM3>(qx, qy, qz, rec + 2 * (jj + 0));
                    const double d1 = dist2_sk<DIM3>(qx, qy, qz, rec + 2 * (jj + 1));
                    const double d2 = dist2_sk<DIM3>(qx, qy, qz, rec + 2 * (jj + 2));
                    const double d3 = dist2_sk<DIM3>(qx, qy, qz, rec + 2 * (jj + 3));
                    const double dmin = fmin(fmin(d0, d1), fmin(d2, d3));
                    if (dmin < top.worst) {
                        const int32_t j = int32_t(c0) + jj;
                        if (d0 < top.worst) top.insert(d0, j);
                        if (d1 < top.worst) top.insert(d1, j + 1);
                        if (d2 < top.worst) top.insert(d2, j + 2);
                        if (d3 < top.worst) top.insert(d3, j + 3);
                    }
                }
            } else {
                // diagonal / ragged tile: only predecessors j < i count
                int lim = 0;
                if (live) {
                    const int64_t v = ORDERED ? (i < cand_cap ? i : cand_cap) - c0 : int64_t(jn);
                    lim = v < 0 ? 0 : (v > jn ? jn : int(v));
                }
                for (int jj = 0; jj < jn; ++jj) {
                    const double d = dist2_sk<DIM3>(qx, qy, qz, rec + 2 * jj);
                    if (jj < lim && d < top.worst) top.insert(d, int32_t(c0) + jj);
                }
            }
            __syncthreads();  // stage s is free again
            if (tid == 0 && c + NSTAGE < nct) issue(c + NSTAGE);
        }

        if (live && i >= out_lo) {  // rows below out_lo share the first tile but are not wanted (and may not exist)
            int32_t *row = out + i * m;
            for (int k = 0; k < m; ++k) row[k] = k < top.cnt ? top.ids[k * TQ] : -1;
        }
        __syncthreads();  // s_tile and the lists are reused by the next tile
    }
}

__global__ void fill_i32_kernel(int32_t *p, int64_t count, int32_t v)
{
    for (int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; k < count;
         k += int64_t(gridDim.x) * blockDim.x)
        p[k] = v;
}

inline size_t smem_bytes(int m)
{
    return size_t(NSTAGE) * TC * sizeof(double4) + size_t(m) * TQ * (sizeof(double) + sizeof(int32_t));
}

}  // namespace nngp_knn

static cudaError_t launch_knn(nngp_handle *h, bool ordered, int m, int tile_offset, int tile_stride,
                              int32_t *table, cudaStream_t stream, int64_t n_rows = -1, int64_t first_row = 0,
                              int64_t cand_cap = INT64_MAX)
{
    using namespace nngp_knn;
    const int64_t n = n_rows >= 0 ? n_rows : h->n;  // rows >= n are neither queries nor candidates
    const int ntiles = int((n + TQ - 1) / TQ);
    cudaError_t e;
    if (tile_stride > 1) {
        fill_i32_kernel<<<h->num_sms * 4, 256, 0, stream>>>(table, n * int64_t(m), NNGP_ROW_UNSET);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        ++h->launches;
    }
    if ((e = cudaMemsetAsync(h->d_tile_counter, 0, sizeof(unsigned int), stream)) != cudaSuccess) return e;
    const size_t smem = smem_bytes(m);
    auto kern = ordered ? (h->D == 3 ? knn_ordered_kernel<true, true> : knn_ordered_kernel<false, true>)
                        : (h->D == 3 ? knn_ordered_kernel<true, false> : knn_ordered_kernel<false, false>);
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
        return e;
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TQ, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    const int first_tile = int(first_row / TQ);
    const int my_tiles = (ntiles - first_tile - tile_offset + tile_stride - 1) / tile_stride;
    int grid = h->num_sms * per_sm;
    if (grid > my_tiles) grid = my_tiles > 0 ? my_tiles : 1;
    kern<<<grid, TQ, smem, stream>>>(h->pts, n, m, ntiles, tile_offset, tile_stride, first_tile, first_row, cand_cap,
                                     h->d_tile_counter, table);
    ++h->launches;
    return cudaGetLastError();
}

cudaError_t launch_knn_ordered(nngp_handle *h, int m, int tile_offset, int tile_stride, cudaStream_t stream)
{
    return launch_knn(h, true, m, tile_offset, tile_stride, h->nbr, stream);
}

cudaError_t launch_knn_brute_rows(nngp_handle *h, int m, int64_t first_row, int64_t n_rows, int64_t cand_cap,
                                  int32_t *d_table, cudaStream_t stream)
{
    return launch_knn(h, true, m, 0, 1, d_table, stream, n_rows, first_row, cand_cap);
}

cudaError_t launch_knn_plain(nngp_handle *h, int k, int32_t *d_table, cudaStream_t stream)
{
    return launch_knn(h, false, k, 0, 1, d_table, stream);
}
