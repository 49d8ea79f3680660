// knn_grid.cu -- stage 1 of the NNGP hot path in sub-quadratic time: the ordered k-nearest-neighbour
// search over a uniform cell grid, bit-identical to the brute-force kernel (knn_ordered.cu).
//
// Takes over _make_s_neighbor_sets, pyNNGP/nngp.py:49-62 (a KDTree rebuild per location upstream).
// The answer for row i is the first min(m, i) predecessors j < i in the total order (d2, j), d2 the
// scikit-learn fp64 squared distance (sklearn/metrics/_dist_metrics.pxd.tp:39-49: per-dimension
// products and sums rounded separately).  The brute-force kernel enumerates all j < i; this one
// proves most of them irrelevant:
//   - the ordering is cut into levels [a, b), b = 2a (rows below `brute_rows` stay with the
//     brute-force kernel).  A level's grid holds the points 0 .. b-1 (every possible predecessor of
//     its queries) with a cell side chosen so that a cell holds ~lambda of the a points every query
//     of the level may use: the search radius follows the predecessor density, which grows with i;
//   - per level: cell histogram (atomics) -> single-block exclusive scan -> scatter into a
//     cell-sorted copy of the records {x, y, z, j} (32 bytes, contiguous per cell and per row of
//     cells) + a cell-sorted list of the level's queries, so consecutive threads search the same
//     cells and their loads hit L1;
//   - one thread per query walks Chebyshev rings of cells around its own cell.  After ring r every
//     unvisited point lies beyond a face of the (2r+1)-cell box, i.e. at distance >= `bound` (faces
//     on the domain boundary do not count: nothing lies beyond them).  The walk stops once the m-th
//     best d2 is below (bound * (1 - 1e-9) - slack)^2, `slack` covering the rounding of the cell
//     assignment -- conservative, so stopping never changes the result; without a stop the walk
//     ends when the box covers the grid (= brute force over the level).  Candidates are compared in
//     the full (d2, j) order because cells are not visited in index order.
// Work: ~3^D * 2 * lambda distance evaluations per query instead of i: 1e8 pair evaluations at
// n = 1e6, m = 15 (brute force: 5e11).  Clustered data only costs time, never exactness; when the
// top-level histogram predicts more work than brute force the caller is told to use that instead.
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <type_traits>

#include "nngp_common.cuh"
#include "peer_exchange.cuh"

namespace nngp_grid {

struct GridSpec {
    double lo[3];     // bounding box minimum
    double h[3];      // cell side per dimension
    double inv_h[3];  // 1 / h (0 for a collapsed dimension: every point in cell 0)
    int G[3];         // cells per dimension (>= 1)
    int ncell;
    double slack;     // absolute rounding allowance of the cell assignment
};

__device__ __forceinline__ int cell_index(const GridSpec &gs, double x, double y, double z)
{
    const int cx = min(gs.G[0] - 1, int((x - gs.lo[0]) * gs.inv_h[0]));
    const int cy = min(gs.G[1] - 1, int((y - gs.lo[1]) * gs.inv_h[1]));
    const int cz = min(gs.G[2] - 1, int((z - gs.lo[2]) * gs.inv_h[2]));
    return (cz * gs.G[1] + cy) * gs.G[0] + cx;
}

// cell of every point 0 .. N-1; histogram of the candidate points 0 .. ncand-1 and of the level's
// queries [qlo, qhi)
template <bool DIM3>
__global__ void cell_count_kernel(const double4 *__restrict__ pts, int N, int ncand, GridSpec gs, int qlo, int qhi,
                                  int *__restrict__ cell_of, int *counts, int *qcounts)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const double4 p = pts[i];
        const int c = cell_index(gs, p.x, p.y, DIM3 ? p.z : gs.lo[2]);
        cell_of[i] = c;
        if (i < ncand) atomicAdd(counts + c, 1);
        if (i >= qlo && i < qhi) atomicAdd(qcounts + c, 1);
    }
}

// Exclusive scan of one count array per block (blockIdx.x = 0: all points, 1: queries): starts[c],
// starts[ncell] = total, cursor[c] = starts[c] (the scatter's running position).  Block 0 also returns
// sum of count^2, the grid search's work estimate.
__global__ void __launch_bounds__(1024) scan_cells_kernel(const int *counts, int *starts, int *cursor,
                                                          const int *qcounts, int *qstarts, int *qcursor,
                                                          int ncell, double *sumsq)
{
    const int *in = blockIdx.x == 0 ? counts : qcounts;
    int *st = blockIdx.x == 0 ? starts : qstarts;
    int *cu = blockIdx.x == 0 ? cursor : qcursor;
    __shared__ int warp_tot[32];
    __shared__ int chunk_tot;
    __shared__ double red[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int running = 0;
    double sq = 0.0;
    constexpr int IPT = 4;  // cells per thread and round (consecutive lanes read consecutive 16-byte groups; 16 per thread made the
                            // accesses 64-byte strided and the kernel 1.6x slower)
    // the next round's counts are loaded before this round's barriers (a round is otherwise one L2 round trip long)
    int vnext[IPT];
#pragma unroll
    for (int k = 0; k < IPT; ++k) vnext[k] = tid * IPT + k < ncell ? in[tid * IPT + k] : 0;
    for (int base = 0; base < ncell; base += 1024 * IPT) {
        const int k0 = base + tid * IPT;
        int v[IPT];
        int s = 0;
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            v[k] = vnext[k];
            sq += double(v[k]) * double(v[k]);
            s += v[k];
        }
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const int kn = k0 + 1024 * IPT + k;
            vnext[k] = kn < ncell ? in[kn] : 0;
        }
        int incl = s;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const int wt = warp_tot[lane];
            int wi = wt;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, off);
                if (lane >= off) wi += t;
            }
            warp_tot[lane] = wi - wt;  // exclusive prefix of the warp totals
            if (lane == 31) chunk_tot = wi;  // block total of this chunk
        }
        __syncthreads();
        int pre = running + warp_tot[wid] + incl - s;
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            if (k0 + k < ncell) { st[k0 + k] = pre; cu[k0 + k] = pre; }
            pre += v[k];
        }
        running += chunk_tot;
        __syncthreads();
    }
    if (tid == 0) st[ncell] = running;
    if (blockIdx.x == 0 && sumsq) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
        if (lane == 0) red[wid] = sq;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < 32; ++w) t += red[w];
            *sumsq = t;
        }
    }
}

// records and query ids into cell order (the order inside a cell is arbitrary: the search compares
// (d2, j), so it does not matter)
__global__ void scatter_kernel(const double4 *__restrict__ pts, int N, int ncand, const int *__restrict__ cell_of,
                               int qlo, int qhi, int *cursor, int *qcursor, double4 *__restrict__ sorted,
                               int *__restrict__ qlist)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const int c = cell_of[i];
        if (i < ncand) {
            double4 p = pts[i];
            p.w = __hiloint2double(0, i);  // the record's y value is not needed here: carry the index
            sorted[atomicAdd(cursor + c, 1)] = p;
        }
        if (i >= qlo && i < qhi) qlist[atomicAdd(qcursor + c, 1)] = i;
    }
}

// ---- the query kernel: SW lanes per query, warp-level top-m selection ----------------------------------------
// A query is searched by a sub-warp of SW lanes (SW = 8 / 16 / 32 >= m; a warp carries 32 / SW queries, neighbours
// in the cell-sorted query list, so its sub-warps walk the same cells).  The m best (d2, j) found so far live in
// REGISTERS, one entry per lane: lane k of the sub-warp holds the k-th best in the total order (d2, j), empty
// entries hold (+inf, INT_MAX) and sort last.  A span of records (one row of cells) is read SW at a time --
// consecutive lanes, consecutive 32-byte records -- every lane evaluates the scikit-learn squared distance of its
// candidate, and a ballot collects the candidates that beat the sub-warp's current m-th best.  Each survivor is
// then inserted by the whole sub-warp: its (d2, j) is broadcast from the lane that holds it, every list lane
// compares its own entry against it, the population count of that ballot is the insertion position, and the tail
// of the list moves up one lane with __shfl_up_sync (the old m-th best falls off the end).  Nothing is kept in
// shared memory and no lane loops over the list.  (d2, j) is compared with FP64 set-predicate instructions: the kernel
// is bound by its issue slots (80 % active, FP64 pipe 3 %), and a 64-bit integer compare costs two to three more
// instructions than a DSETP.
// The threshold the ballot uses is refreshed once per SW candidates, not per insertion: a stale threshold only
// lets a few candidates through that land beyond position m - 1, where they are dropped.
constexpr int QTHREADS = 256;  // threads per block of the query kernel
constexpr int32_t kNoId = 0x7fffffff;

struct Cand {
    double key;  // d2 (never NaN: the coordinates are finite)
    int32_t id;
};
__device__ __forceinline__ bool cand_less(double ka, int32_t ia, double kb, int32_t ib)
{
    return ka < kb || (ka == kb && ia < ib);
}

// ORDERED: candidates are the predecessors j < i.  !ORDERED: every grid point, the query included
// (the plain k-NN behind the reference's `ws`, nngp.py:45-47).
template <bool DIM3, bool ORDERED, int SW>
__global__ void __launch_bounds__(QTHREADS) knn_grid_query_kernel(const double4 *__restrict__ pts,
                                                                  const double4 *__restrict__ sorted,
                                                                  const int *__restrict__ starts,
                                                                  const int *__restrict__ cell_of,
                                                                  const int *__restrict__ qlist, int nq, GridSpec gs, int m,
                                                                  int cand_cap, int32_t *__restrict__ out)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr unsigned SUBMASK = 0xffffffffu >> (32 - SW);
    const double kInf = INFINITY;
    const int lane = threadIdx.x & 31;
    const int sl = lane % SW;             // lane inside the sub-warp = list position it holds
    const int sshift = lane - sl;         // first lane of the sub-warp
    const int t = (blockIdx.x * (QTHREADS / 32) + (threadIdx.x >> 5)) * (32 / SW) + lane / SW;
    // a sub-warp without a query walks along with an empty box (whole warps leave)
    if ((blockIdx.x * (QTHREADS / 32) + (threadIdx.x >> 5)) * (32 / SW) >= nq) return;
    const bool has_q = t < nq;
    const int i = has_q ? qlist[t] : 0;
    const double4 q = pts[i];
    const double qx = q.x, qy = q.y, qz = q.z;
    const int Gx = gs.G[0], Gy = gs.G[1], Gz = gs.G[2];
    const int jlim = i < cand_cap ? i : cand_cap;  // ORDERED: candidates are j < min(i, cand_cap)
    int c = cell_of[i];  // as assigned by cell_count_kernel: query and candidates use one assignment
    const int cx = c % Gx;
    c /= Gx;
    const int cy = c % Gy, cz = c / Gy;

    double key = kInf;  // this lane's list entry
    int32_t id = kNoId;
    double worst = kInf;  // the sub-warp's m-th best (the acceptance threshold), refreshed after a batch that inserted
    int32_t worst_id = kNoId;

    const double2 *rec = reinterpret_cast<const double2 *>(sorted);
    // scans records [s, e) of the sub-warp's query (s >= e: nothing); sub-warps of a warp run in lock step
    auto scan_span = [&](int s, int e) {
        for (int base = s; __any_sync(FULL, base < e); base += SW) {
            const int k = base + sl;
            Cand cd{kInf, kNoId};
            bool pass = false;
            if (k < e) {
                const double2 a = __ldg(rec + 2 * k), b = __ldg(rec + 2 * k + 1);
                cd.id = __double2loint(b.y);
                // scikit-learn's order of operations, no contraction (bit-exact with knn_ordered.cu)
                const double dx = qx - a.x, dy = qy - a.y;
                double d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                if (DIM3) {
                    const double dz = qz - b.x;
                    d = __dadd_rn(d, __dmul_rn(dz, dz));
                }
                cd.key = d;
                pass = (!ORDERED || cd.id < jlim) && cand_less(cd.key, cd.id, worst, worst_id);
            }
            unsigned todo = (__ballot_sync(FULL, pass) >> sshift) & SUBMASK;  // this sub-warp's survivors
            if (!__any_sync(FULL, todo != 0)) continue;  // nothing to insert anywhere in the warp: the thresholds stand
            while (__any_sync(FULL, todo != 0)) {
                const bool act = todo != 0;
                const int src = act ? __ffs(todo) - 1 : 0;
                const double ck = __shfl_sync(FULL, cd.key, src, SW);
                const int32_t cj = __shfl_sync(FULL, cd.id, src, SW);
                // list entries that sort before the candidate form a prefix: its length is the position
                const bool before = sl < m && cand_less(key, id, ck, cj);
                const int pos = __popc((__ballot_sync(FULL, before) >> sshift) & SUBMASK);
                const double kup = __shfl_up_sync(FULL, key, 1, SW);
                const int32_t iup = __shfl_up_sync(FULL, id, 1, SW);
                if (act) {
                    if (sl == pos) { key = ck; id = cj; }
                    else if (sl > pos) { key = kup; id = iup; }
                }
                todo &= todo - 1;
            }
            worst = __shfl_sync(FULL, key, m - 1, SW);
            worst_id = __shfl_sync(FULL, id, m - 1, SW);
        }
    };

    bool done = !has_q;
    for (int r = 1; !__all_sync(FULL, done); ++r) {
        const int x0 = max(cx - r, 0), x1 = min(cx + r, Gx - 1);
        const int y0 = max(cy - r, 0), y1 = min(cy + r, Gy - 1);
        const int z0 = max(cz - r, 0), z1 = min(cz + r, Gz - 1);
        // lock step over the (2r+1)^2 rows of cells of the box: rows outside a sub-warp's own box are empty spans
        for (int dz = DIM3 ? -r : 0; dz <= (DIM3 ? r : 0); ++dz) {
            const int zz = cz + dz;
            const bool zin = !done && zz >= z0 && zz <= z1;
            const bool zshell = dz == r || dz == -r;
            for (int dy = -r; dy <= r; ++dy) {
                const int yy = cy + dy;
                const bool in = zin && yy >= y0 && yy <= y1;
                const int rowbase = (zz * Gy + yy) * Gx;
                if (r == 1 || (DIM3 && zshell) || dy == r || dy == -r) {
                    // a row of cells new to this ring: one contiguous span of records
                    int s = 0, e = 0;
                    if (in) { s = starts[rowbase + x0]; e = starts[rowbase + x1 + 1]; }
                    scan_span(s, e);
                } else {
                    int s = 0, e = 0;
                    if (in && cx - r >= 0) { s = starts[rowbase + cx - r]; e = starts[rowbase + cx - r + 1]; }
                    scan_span(s, e);
                    s = 0; e = 0;
                    if (in && cx + r <= Gx - 1) { s = starts[rowbase + cx + r]; e = starts[rowbase + cx + r + 1]; }
                    scan_span(s, e);
                }
            }
        }
        if (!done) {
            if (x0 == 0 && x1 == Gx - 1 && y0 == 0 && y1 == Gy - 1 && z0 == 0 && z1 == Gz - 1) done = true;
            else if (worst_id != kNoId) {  // the list is full
                // distance from the query to the nearest face of the visited box that has cells behind it
                double bound = INFINITY;
                if (cx - r > 0) bound = fmin(bound, qx - (gs.lo[0] + double(cx - r) * gs.h[0]));
                if (cx + r < Gx - 1) bound = fmin(bound, (gs.lo[0] + double(cx + r + 1) * gs.h[0]) - qx);
                if (cy - r > 0) bound = fmin(bound, qy - (gs.lo[1] + double(cy - r) * gs.h[1]));
                if (cy + r < Gy - 1) bound = fmin(bound, (gs.lo[1] + double(cy + r + 1) * gs.h[1]) - qy);
                if (cz - r > 0) bound = fmin(bound, qz - (gs.lo[2] + double(cz - r) * gs.h[2]));
                if (cz + r < Gz - 1) bound = fmin(bound, (gs.lo[2] + double(cz + r + 1) * gs.h[2]) - qz);
                const double bs = bound * (1.0 - 1e-9) - gs.slack;
                if (bs > 0.0 && worst < bs * bs) done = true;
            }
        }
    }

    if (has_q && sl < m) out[int64_t(i) * m + sl] = id == kNoId ? -1 : id;
}

__global__ void fill_rows_kernel(int32_t *p, int64_t count, int32_t v)
{
    for (int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; k < count;
         k += int64_t(gridDim.x) * blockDim.x)
        p[k] = v;
}

// Cell grid for `nref` reference points inside the handle's bounding box with ~lambda of them per
// cell.  Dimensions whose extent is below the cell side collapse to one cell.
GridSpec make_grid(const nngp_handle *h, double nref, double lambda, int64_t max_cells)
{
    GridSpec gs{};
    double ext[3];
    bool active[3];
    double mag = 0.0;
    for (int d = 0; d < 3; ++d) {
        gs.lo[d] = d < h->D ? h->bb_lo[d] : 0.0;
        ext[d] = d < h->D ? h->bb_hi[d] - h->bb_lo[d] : 0.0;
        active[d] = ext[d] > 0.0;
        if (d < h->D) mag = std::max(mag, std::max(fabs(h->bb_lo[d]), fabs(h->bb_hi[d])));
    }
    double side = 0.0;
    for (int pass = 0; pass < 4; ++pass) {
        int deff = 0;
        double vol = 1.0;
        for (int d = 0; d < 3; ++d)
            if (active[d]) { ++deff; vol *= ext[d]; }
        if (deff == 0) break;
        side = pow(lambda * vol / std::max(nref, 1.0), 1.0 / deff);
        bool changed = false;
        for (int d = 0; d < 3; ++d)
            if (active[d] && !(ext[d] >= side)) { active[d] = false; changed = true; }
        if (!changed) break;
    }
    for (;;) {
        int64_t total = 1;
        for (int d = 0; d < 3; ++d) {
            gs.G[d] = 1;
            if (active[d] && side > 0.0) {
                const double g = floor(ext[d] / side);
                gs.G[d] = g < 1.0 ? 1 : (g > 1048576.0 ? 1048576 : int(g));
            }
            total *= gs.G[d];
        }
        if (total <= max_cells) { gs.ncell = int(total); break; }
        side *= 1.26;
    }
    for (int d = 0; d < 3; ++d) {
        if (gs.G[d] > 1) {
            gs.h[d] = ext[d] / gs.G[d];
            gs.inv_h[d] = gs.G[d] / ext[d];
        } else {
            gs.h[d] = 0.0;  // one cell: no interior faces, the side is never used
            gs.inv_h[d] = 0.0;
        }
    }
    gs.slack = 1e-14 * mag;
    return gs;
}

// handle-owned (scratch_get): a build costs no cudaMalloc / cudaFree after the first one
struct Scratch {
    double4 *sorted = nullptr;
    int *cell_of = nullptr, *qlist = nullptr;
    int *cells = nullptr;  // 6 arrays of (cap_cells + 1)
    double *sumsq = nullptr;
};

}  // namespace nngp_grid

cudaError_t launch_fill_i32(nngp_handle *h, int32_t *p, int64_t count, int32_t v, cudaStream_t stream)
{
    if (count <= 0) return cudaSuccess;
    nngp_grid::fill_rows_kernel<<<h->num_sms * 4, 256, 0, stream>>>(p, count, v);
    ++h->launches;
    return cudaGetLastError();
}

// Fills rows [row_lo, row_hi) of `table` (n x m) with the grid search; rows below brute_rows come
// from the brute-force kernel.  ordered = false: one level over all n points, the query itself
// included.  *used = 0 (and nothing computed) when the data does not suit a grid and force == 0:
// non-finite coordinates, or a histogram that predicts more work than brute force.
cudaError_t launch_knn_grid(nngp_handle *h, bool ordered, int m, int64_t row_lo, int64_t row_hi, int64_t cand_cap,
                            int32_t *table, bool window, cudaStream_t stream, int force, int *used)
{
    using namespace nngp_grid;
    *used = 0;
    const int64_t n = h->n;
    if (!h->bb_finite) return cudaSuccess;
    cudaError_t e;
#define GRID_TRY(call)                      \
    do {                                    \
        e = (call);                         \
        if (e != cudaSuccess) return e;     \
    } while (0)

    // rows [0, T0) by brute force (few predecessors: a grid has nothing to prune); ordered only
    int64_t T0 = 0;
    if (ordered) {
        T0 = std::max<int64_t>(h->knn_brute_rows, 4 * int64_t(m));
        T0 = (T0 + NNGP_KNN_TILE - 1) / NNGP_KNN_TILE * NNGP_KNN_TILE;
        if (T0 > n) T0 = n;
    }
    const int deff = std::max(1, (h->bb_hi[0] > h->bb_lo[0]) + (h->D > 1 && h->bb_hi[1] > h->bb_lo[1]) +
                                     (h->D > 2 && h->bb_hi[2] > h->bb_lo[2]));
    const double ball = deff == 1 ? 2.0 : deff == 2 ? 3.141592653589793 : 4.1887902047863905;
    const double lambda = std::max(1.0, h->knn_lambda_scale * (m + 2.0 * sqrt(double(m))) / ball);
    const int64_t max_cells = std::min<int64_t>(std::max<int64_t>(8 * n, 1024), int64_t(1) << 28);  // reached only by refinement

    // levels, top (largest) first; the top level's histogram decides whether a grid pays off
    struct Level { int64_t a, b; };
    Level lev[40];
    int nlev = 0;
    if (ordered) {
        // (the lowest level runs from the brute-force rows up to 4096: a few thousand queries over a few thousand
        // points cost microseconds on any grid, a level's four launches do not)
        for (int64_t a = T0; a < n;) {
            const int64_t b = std::min<int64_t>(std::max<int64_t>(2 * a, a == T0 ? 4096 : 0), n);
            lev[nlev++] = Level{a, b};
            a = b;
        }
    } else {
        lev[nlev++] = Level{0, n};
    }

    bool any_level = false;
    for (int l = nlev - 1; l >= 0; --l)
        if (std::max(lev[l].a, row_lo) < std::min(lev[l].b, row_hi)) any_level = true;

    const bool dim3 = h->D == 3;
    // lanes per query: the smallest sub-warp that holds the m list entries
    const int sw = m <= 8 ? 8 : m <= 16 ? 16 : 32;
    auto pick = [&](auto dim3_c, auto ordered_c) {
        constexpr bool D3 = decltype(dim3_c)::value, ORD = decltype(ordered_c)::value;
        return sw == 8 ? knn_grid_query_kernel<D3, ORD, 8> : sw == 16 ? knn_grid_query_kernel<D3, ORD, 16> : knn_grid_query_kernel<D3, ORD, 32>;
    };
    auto qkern = ordered ? (dim3 ? pick(std::true_type{}, std::true_type{}) : pick(std::false_type{}, std::true_type{}))
                         : (dim3 ? pick(std::true_type{}, std::false_type{}) : pick(std::false_type{}, std::false_type{}));
    const bool partial = !window && (row_lo > 0 || row_hi < n);  // a window table has no rows outside [row_lo, row_hi)
    bool filled = false;

    // Clustered data: when the top level's histogram predicts too much work, the cells are refined (a
    // quarter of the occupancy per attempt: dense regions get small cells, sparse regions cost a few more
    // rings of mostly empty cells) before brute force is considered.
    double lam_use = lambda;
    for (int attempt = 0;; ++attempt) {
    Scratch sc;
    bool rejected = false;
    int cap_cells = 0;
    if (any_level) {
        for (int l = 0; l < nlev; ++l)
            cap_cells = std::max(cap_cells, make_grid(h, double(ordered ? std::min(lev[l].a, cand_cap) : n), lam_use, max_cells).ncell);
        GRID_TRY(scratch_get(h, 0, sizeof(double4) * size_t(n), reinterpret_cast<void **>(&sc.sorted)));
        GRID_TRY(scratch_get(h, 1, sizeof(int) * size_t(n), reinterpret_cast<void **>(&sc.cell_of)));
        GRID_TRY(scratch_get(h, 2, sizeof(int) * size_t(n), reinterpret_cast<void **>(&sc.qlist)));
        GRID_TRY(scratch_get(h, 3, sizeof(int) * 6 * size_t(cap_cells + 1), reinterpret_cast<void **>(&sc.cells)));
        GRID_TRY(scratch_get(h, 4, sizeof(double), reinterpret_cast<void **>(&sc.sumsq)));
    }
    int *counts = sc.cells, *starts = counts + (cap_cells + 1), *cursor = starts + (cap_cells + 1);
    int *qcounts = cursor + (cap_cells + 1), *qstarts = qcounts + (cap_cells + 1), *qcursor = qstarts + (cap_cells + 1);

    bool first = true;
    for (int l = nlev - 1; l >= 0; --l) {
        const int64_t qlo = std::max(lev[l].a, row_lo), qhi = std::min(lev[l].b, row_hi);
        if (qlo >= qhi) continue;
        const int N = int(lev[l].b);
        const int ncand = int(std::min<int64_t>(N, cand_cap));
        const GridSpec gs = make_grid(h, double(ordered ? std::min(lev[l].a, cand_cap) : n), lam_use, max_cells);
        GRID_TRY(cudaMemsetAsync(counts, 0, sizeof(int) * size_t(gs.ncell + 1), stream));
        GRID_TRY(cudaMemsetAsync(qcounts, 0, sizeof(int) * size_t(gs.ncell + 1), stream));
        const int sgrid = int(std::min<int64_t>((N + 255) / 256, int64_t(h->num_sms) * 16));
        if (dim3)
            cell_count_kernel<true><<<sgrid, 256, 0, stream>>>(h->pts, N, ncand, gs, int(qlo), int(qhi), sc.cell_of, counts, qcounts);
        else
            cell_count_kernel<false><<<sgrid, 256, 0, stream>>>(h->pts, N, ncand, gs, int(qlo), int(qhi), sc.cell_of, counts, qcounts);
        GRID_TRY(cudaGetLastError());
        scan_cells_kernel<<<2, 1024, 0, stream>>>(counts, starts, cursor, qcounts, qstarts, qcursor, gs.ncell, sc.sumsq);
        GRID_TRY(cudaGetLastError());
        h->launches += 2;
        if (first) {
            first = false;
            if (!force) {
                // predicted pair evaluations: every query scans ~3^D cells of its own cell's occupancy
                double sumsq = 0.0;
                GRID_TRY(cudaMemcpyAsync(&sumsq, sc.sumsq, sizeof(double), cudaMemcpyDeviceToHost, stream));
                GRID_TRY(cudaStreamSynchronize(stream));
                const double est = pow(3.0, deff) * sumsq;
                if (est > double(ncand) * double(ncand) / 16.0) { rejected = true; break; }  // nothing written yet
            }
        }
        if (partial && !filled) {
            fill_rows_kernel<<<h->num_sms * 4, 256, 0, stream>>>(table, n * int64_t(m), NNGP_ROW_UNSET);
            GRID_TRY(cudaGetLastError());
            ++h->launches;
            filled = true;
        }
        scatter_kernel<<<sgrid, 256, 0, stream>>>(h->pts, N, ncand, sc.cell_of, int(qlo), int(qhi), cursor, qcursor, sc.sorted,
                                                  sc.qlist);
        GRID_TRY(cudaGetLastError());
        const int nq = int(qhi - qlo);
        const int qpb = (QTHREADS / 32) * (32 / sw);  // queries per block
        qkern<<<(nq + qpb - 1) / qpb, QTHREADS, 0, stream>>>(h->pts, sc.sorted, starts, sc.cell_of, sc.qlist, nq, gs, m,
                                                             int(std::min<int64_t>(cand_cap, INT32_MAX)), table);
        GRID_TRY(cudaGetLastError());
        h->launches += 2;
    }
    if (!rejected) break;
    if (attempt == 4 || int64_t(cap_cells) * 2 > max_cells) return cudaSuccess;  // *used stays 0: brute force
    lam_use *= 0.25;
    }  // attempts
    if (partial && !filled) {
        fill_rows_kernel<<<h->num_sms * 4, 256, 0, stream>>>(table, n * int64_t(m), NNGP_ROW_UNSET);
        GRID_TRY(cudaGetLastError());
        ++h->launches;
    }
    if (ordered && row_lo < T0 && T0 > 0)
        GRID_TRY(launch_knn_brute_rows(h, m, row_lo, std::min(T0, row_hi), cand_cap, table, stream));
    GRID_TRY(cudaStreamSynchronize(stream));
#undef GRID_TRY
    *used = 1;
    return cudaSuccess;
}

// A rank with an empty shard still takes part in the statistics exchange: one block per parameter vector
// publishes zeros and collects the sum (peer_exchange.cuh).
__global__ void __launch_bounds__(32) peer_zero_kernel(PeerExchange px, EvalArgs a)
{
    __shared__ double sh[3 + 2 * 24];
    double tot[3] = {0.0, 0.0, 0.0};
    if (px.world > 1) peer_allreduce3(px, blockIdx.x, 0.0, 0.0, 0.0, sh, tot);
    if (threadIdx.x == 0) publish_result(a, blockIdx.x, tot);
}

cudaError_t launch_peer_zero(nngp_handle *, const PeerExchange &px, int K, double *d_out, uint4 *hout, unsigned int seq,
                             cudaStream_t stream)
{
    EvalArgs a{};
    a.out = d_out; a.hout = hout; a.seq = seq;
    peer_zero_kernel<<<K, 32, 0, stream>>>(px, a);
    return cudaGetLastError();
}
