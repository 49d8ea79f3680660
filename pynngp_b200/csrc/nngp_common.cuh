// nngp_common.cuh -- shared declarations of libnngp_b200.so (sm_100a only).
//
// Data layout in HBM (per handle, one B200):
//   pts   n x double4 {x, y, z, yval}   32-byte records: one L2 sector per neighbour gather,
//                                       and a contiguous byte range per candidate tile for the
//                                       k-NN's bulk (TMA 1-D) copies.  z = eps2 (or 0) when D < 3, y = 0 when D < 2.
//   eps2  n x double (optional, D = 3)  per-observation variance added to the diagonal; for D < 3 it
//                                       rides in the record's unused z slot instead.
//   nbr   n x m int32, row-major        neighbour table, -1 padded, valid entries first.
// Coordinates and y are replicated on every GPU; a handle evaluates rows [lo, hi) only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/nngp_b200.h"

// Cross-GPU exchange of the K x 3 statistics over NVLink peer memory (buffers of other processes mapped
// through CUDA IPC, buffers of other devices of this process through cudaDeviceEnablePeerAccess).  Every
// rank owns one exchange buffer of 16-byte lines
//   lines [2][NNGP_MAX_PEERS][K_cap][3] x {lo32(value), gen32, hi32(value), gen32};
// the leading index is the parity of the exchange generation, the second the WRITING rank.  A line is
// written by ONE 16-byte store and carries its own generation stamp beside each half of the value, so the
// reader needs neither a separate flag nor a fence: it polls the line until both stamps match.
constexpr int NNGP_MAX_PEERS = 8;
struct PeerExchange {
    int world = 0, rank = 0, K_cap = 0;
    unsigned int gen = 0;                        // this launch's generation (>= 1), same on all ranks
    uint4 *lines[NNGP_MAX_PEERS] = {};           // rank r's buffer as mapped in this process
};

struct nngp_handle {
    int device = 0;
    int dtype = NNGP_F64;
    int num_sms = 0;
    cudaStream_t stream = nullptr;

    int64_t n = 0;
    int D = 0;
    int m = 0;
    int64_t lo = 0, hi = 0;
    double bb_lo[3] = {0, 0, 0}, bb_hi[3] = {0, 0, 0};  // bounding box of the coordinates (set_data)
    bool bb_finite = false;                              // every coordinate finite
    double knn_lambda_scale = 0.5;                       // grid k-NN: cell occupancy multiplier
    int knn_used_grid = 0;                               // last stage-1 build went through the grid
    int64_t knn_brute_rows = 128;                        // grid k-NN: rows below this use brute force

    double4 *pts = nullptr;
    double *d_ystage = nullptr;  // n doubles: landing buffer of nngp_set_y (allocated on first use)
    double *eps2 = nullptr;
    // neighbour table: rows [nbr_row0, nbr_row0 + nbr_rows) of the n x m table (all n rows unless the table
    // was built for the shard only, nngp_build_neighbors_shard); kernels index it through nbr_base()
    int32_t *nbr = nullptr;
    int64_t nbr_row0 = 0, nbr_rows = 0;
    bool has_nbr = false;
    int32_t *nbr_base() const { return nbr - nbr_row0 * m; }
    bool holds_rows(int64_t i0, int64_t i1) const { return i0 >= nbr_row0 && i1 <= nbr_row0 + nbr_rows; }

    // evaluation scratch (grown on demand)
    int K_cap = 0;
    int grid_cap = 0;
    double *d_params = nullptr;    // K_cap x 4
    double *d_out = nullptr;       // K_cap x 3
    double *d_partials = nullptr;  // K_cap x grid_cap x 3
    unsigned int *d_counters = nullptr;  // K_cap tickets for the last-block reduction
    unsigned int *d_tile_counter = nullptr;
    double *d_exp2tab = nullptr;   // 2^(j/2048), j < 2048, then 16 copies of 2^(j/256), j < 256 (the device's shared table, not owned)
    double *h_stage = nullptr;     // pinned: K_cap x 4 parameter staging (K > NNGP_PV_MAX only)
    // results of the host-pointer calls land in MAPPED pinned host memory, written by the kernel's last
    // block itself, followed by a sequence stamp per blockIdx.y the host polls: no D2H copy, no stream sync
    uint4 *h_out = nullptr;                  // K_cap x 3 stamped lines (host address == device address under UVA)
    unsigned int seq = 0;                    // stamp of the last host-pointer evaluation
    int32_t *d_viol = nullptr;               // violation counter of nngp_set_neighbors' table check

    // stage-1 scratch of the grid search, kept between builds (grown on demand, freed with the handle)
    void *knn_scratch[5] = {};
    size_t knn_scratch_bytes[5] = {};

    // peer exchange (multi-GPU): own buffer, peers' buffers opened through CUDA IPC
    void *xbuf = nullptr;
    void *peer_base[NNGP_MAX_PEERS] = {};
    PeerExchange px;

    int shape_key = -1, shape_per_sm = 1, shape_lpw = 8;  // cached occupancy of the fused kernel in use

    // a *_device entry point ran on a caller's stream: before a buffer goes back to the pool the whole device is
    // synchronised (what cudaFree did implicitly), not just the handle's stream
    bool foreign_stream_used = false;

    int64_t launches = 0;
    std::string err;
    // optional device-side timing of the evaluation launches (nngp_set_timing): events around the kernel
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // multi-device handle (nngp_create_multi): one sub-handle per device, driven by worker threads
    struct nngp_group *group = nullptr;
};

// Device memory of a handle comes from the device's stream-ordered pool (cudaMallocAsync / cudaFreeAsync on the
// handle's stream, release threshold raised at nngp_create): a freed block goes back to the pool, not to the driver,
// so rebuilding a model -- or the next handle -- reuses it without a cudaMalloc / cudaFree round trip (those cost
// 0.1-1 ms each and cudaFree synchronises the device: 3 ms of a 5.5 ms constructor at n = 1e5, and up to 25 ms when
// large blocks were released just before).  The exchange buffers stay on cudaMalloc (CUDA IPC and peer access are
// not defined for pool memory by default).
template <typename P>
inline cudaError_t dev_malloc_on(nngp_handle *h, P **p, size_t bytes)
{
    return cudaMallocAsync(reinterpret_cast<void **>(p), bytes ? bytes : 1, h->stream);
}
// nothing of the handle's may still be in flight (call before buffers are released or replaced)
inline cudaError_t quiesce(nngp_handle *h)
{
    if (h->foreign_stream_used) {
        h->foreign_stream_used = false;
        return cudaDeviceSynchronize();
    }
    return h->stream ? cudaStreamSynchronize(h->stream) : cudaSuccess;
}
template <typename P>
inline void free_dev_on(nngp_handle *h, P *&p)
{
    if (p) cudaFreeAsync(p, h->stream);
    p = nullptr;
}

// Arguments of the fused covariance + factorisation + reduction kernel.
constexpr int NNGP_PV_MAX = 8;  // parameter vectors that travel in the kernel arguments (no H2D copy)
struct EvalArgs {
    const double4 *pts;
    const double *eps2;   // nullable
    const int32_t *nbr;   // n x m (base of row 0; a handle may hold a window of rows only)
    int64_t lo, hi;       // rows to evaluate
    int m;
    const double *params;     // K x 4 (device), or NULL: the vectors are in pv[] (K <= NNGP_PV_MAX)
    double pv[NNGP_PV_MAX][NNGP_NPARAM];
    int K;                    // parameter vectors of this launch
    const double *exp2tab;    // 2^(j/2048), j < 2048; [2048 + j * 16 + c] = 2^(j/256), j < 256, c < 16
    double *partials;         // gridDim.y x gridDim.x x 3
    unsigned int *counters;   // gridDim.y
    double *out;              // gridDim.y x 3 doubles in device memory, or NULL:
    uint4 *hout;              //   gridDim.y x 3 self-stamped 16-byte lines in MAPPED pinned host memory (ll_store)
    unsigned int seq;         //   the stamp of this evaluation
    // optional per-location outputs (rows lo..hi map to output rows 0..hi-lo); any may be null
    int gather_bypass_l1;     // records >> L2: gather them with cp.async.cg (see loglik_fused.cuh issue_rec)
    int emit;                 // 0: reduction only
    double *B, *F, *CN, *cc, *cs;
    PeerExchange px;          // world > 1: `out` receives the sum over all ranks (see peer_allreduce3)
};

// launchers implemented in the .cu files; each returns the cudaError of the launch.
cudaError_t launch_fused_loglik(nngp_handle *h, int kernel_id, const EvalArgs &a, int K,
                                cudaStream_t stream);
// v (n doubles, device) -> lane `lane` of the records (3 = yval; 2 = eps2 when D < 3)
cudaError_t launch_scatter_lane(nngp_handle *h, const double *d_v, int lane, cudaStream_t stream);
cudaError_t launch_knn_ordered(nngp_handle *h, int m, int tile_offset, int tile_stride,
                               cudaStream_t stream);
cudaError_t launch_knn_plain(nngp_handle *h, int k, int32_t *d_table, cudaStream_t stream);
cudaError_t launch_fill_i32(nngp_handle *h, int32_t *p, int64_t count, int32_t v, cudaStream_t stream);
// brute force for rows [first_row, n_rows) of the ordering (whole query tiles are searched, only these rows are
// written); candidates of row i are j < min(i, cand_cap).  d_table is the base of row 0.
cudaError_t launch_knn_brute_rows(nngp_handle *h, int m, int64_t first_row, int64_t n_rows, int64_t cand_cap,
                                  int32_t *d_table, cudaStream_t stream);
// grid search (knn_grid.cu) for rows [row_lo, row_hi), candidates j < min(i, cand_cap); *used = 0 when
// the data does not suit a grid
// `table` is the base of row 0; window = true: only rows [row_lo, row_hi) exist behind it (nothing else is written)
cudaError_t launch_knn_grid(nngp_handle *h, bool ordered, int m, int64_t row_lo, int64_t row_hi, int64_t cand_cap,
                            int32_t *table, bool window, cudaStream_t stream, int force, int *used);
// counts into *d_viol the rows i in [i0, i1) of `rows` (row i0 first) that hold an entry outside [-1, i) or a
// valid entry after a -1
cudaError_t launch_validate_table(nngp_handle *h, const int32_t *rows, int m, int64_t i0, int64_t i1, int32_t *d_viol,
                                  cudaStream_t stream);
// device memory that stays with the handle between calls (slot 0..4), grown on demand
cudaError_t scratch_get(nngp_handle *h, int slot, size_t bytes, void **p);
// publishes zeros for a rank whose shard is empty (it still takes part in the exchange)
cudaError_t launch_peer_zero(nngp_handle *h, const PeerExchange &px, int K, double *d_out, uint4 *hout, unsigned int seq,
                             cudaStream_t stream);
// device-side packing of the records + bounding box (pack.cu); synchronises `stream`
cudaError_t launch_pack_records(nngp_handle *h, const double *d_coords, const double *d_y, const double *d_eps2,
                                cudaStream_t stream);
cudaError_t launch_fma_peak(nngp_handle *h, int dtype, int iters, double *instr_per_s);
