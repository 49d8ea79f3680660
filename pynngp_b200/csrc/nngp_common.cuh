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

// Cross-GPU exchange of the K x 3 statistics over NVLink peer memory (one process per GPU, buffers
// shared through CUDA IPC).  Every rank owns one exchange buffer:
//   slots [2][NNGP_MAX_PEERS][K_cap * 3] doubles, then flags [2][NNGP_MAX_PEERS][K_cap] generations;
// the leading index is the parity of the exchange generation, the second the WRITING rank.
constexpr int NNGP_MAX_PEERS = 8;
struct PeerExchange {
    int world = 0, rank = 0, K_cap = 0;
    unsigned long long gen = 0;                  // this launch's generation (>= 1), same on all ranks
    double *slots[NNGP_MAX_PEERS] = {};          // rank r's buffer as mapped in this process
    unsigned long long *flags[NNGP_MAX_PEERS] = {};
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
    double knn_lambda_scale = 1.0;                       // grid k-NN: cell occupancy multiplier
    int knn_used_grid = 0;                               // last stage-1 build went through the grid
    int64_t knn_brute_rows = 4096;                       // grid k-NN: rows below this use brute force

    double4 *pts = nullptr;
    double *d_ystage = nullptr;  // n doubles: landing buffer of nngp_set_y (allocated on first use)
    double *eps2 = nullptr;
    int32_t *nbr = nullptr;
    bool has_nbr = false;

    // evaluation scratch (grown on demand)
    int K_cap = 0;
    int grid_cap = 0;
    double *d_params = nullptr;    // K_cap x 4
    double *d_out = nullptr;       // K_cap x 3
    double *d_partials = nullptr;  // K_cap x grid_cap x 3
    unsigned int *d_counters = nullptr;  // K_cap tickets for the last-block reduction
    unsigned int *d_tile_counter = nullptr;
    double *d_exp2tab = nullptr;   // 2^(j/2048), j < 2048 (built once at nngp_create)
    double *h_stage = nullptr;     // pinned: K_cap x (4 + 3)

    // peer exchange (multi-GPU): own buffer, peers' buffers opened through CUDA IPC
    void *xbuf = nullptr;
    void *peer_base[NNGP_MAX_PEERS] = {};
    PeerExchange px;

    int shape_key = -1, shape_per_sm = 1, shape_lpw = 8;  // cached occupancy of the fused kernel in use

    int64_t launches = 0;
    std::string err;
};

// Arguments of the fused covariance + factorisation + reduction kernel.
struct EvalArgs {
    const double4 *pts;
    const double *eps2;   // nullable
    const int32_t *nbr;   // n x m
    int64_t lo, hi;       // rows to evaluate
    int m;
    const double *params;     // K x 4 (device)
    int K;                    // parameter vectors of this launch
    const double *exp2tab;    // 2^(j/2048), j < 2048
    double *partials;         // gridDim.y x gridDim.x x 3
    unsigned int *counters;   // gridDim.y
    double *out;              // gridDim.y x 3
    // optional per-location outputs (rows lo..hi map to output rows 0..hi-lo); any may be null
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
// brute force restricted to the query tiles covering rows [first_row, n_rows) of the ordering;
// candidates of row i are j < min(i, cand_cap)
cudaError_t launch_knn_brute_rows(nngp_handle *h, int m, int64_t first_row, int64_t n_rows, int64_t cand_cap,
                                  int32_t *d_table, cudaStream_t stream);
// grid search (knn_grid.cu) for rows [row_lo, row_hi), candidates j < min(i, cand_cap); *used = 0 when
// the data does not suit a grid
cudaError_t launch_knn_grid(nngp_handle *h, bool ordered, int m, int64_t row_lo, int64_t row_hi, int64_t cand_cap,
                            int32_t *table, cudaStream_t stream, int force, int *used);
// publishes zeros for a rank whose shard is empty (it still takes part in the exchange)
cudaError_t launch_peer_zero(nngp_handle *h, const PeerExchange &px, int K, double *d_out, cudaStream_t stream);
// device-side packing of the records + bounding box (pack.cu); synchronises `stream`
cudaError_t launch_pack_records(nngp_handle *h, const double *d_coords, const double *d_y, const double *d_eps2,
                                cudaStream_t stream);
cudaError_t launch_fma_peak(nngp_handle *h, int dtype, int iters, double *instr_per_s);
