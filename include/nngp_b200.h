/*
 * nngp_b200.h -- C ABI of libnngp_b200.so: the B200 (sm_100a) NNGP likelihood hot path.
 *
 * The reference (bwpriest/pyNNGP) is a pure-Python class with no FFI of its own; each entry point
 * below names the reference symbol (pyNNGP/nngp.py:LINE) whose work it takes over, and
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success and a nonzero NNGP_E* code on failure; the message is
 *     available from nngp_last_error(h) (or nngp_last_error(NULL) for a failed nngp_create).
 *   - a device handle (nngp_create) = one CUDA device = one shard of the ordering (rows [lo, hi) of the
 *     location list).  Multi-GPU, two ways:
 *       (a) one process per GPU, one device handle each; the sum of the 3*K partial statistics over the
 *           ranks is fused into the kernel's tail over NVLink peer memory (nngp_peer_export /
 *           nngp_peer_connect / nngp_loglik_allreduce), or left to the caller's allreduce;
 *       (b) ONE process, a multi-device handle (nngp_create_multi): the same entry points fan out over
 *           one sub-handle and one launcher thread per device, the shard is split evenly over the
 *           devices, and nngp_loglik returns the total -- what a plain caller of the reference's
 *           constructor (pyNNGP/nngp.py:6) gets without torchrun.
 *   - the caller owns every host buffer; the library copies in / out and never keeps host
 *     pointers.  The handle owns device memory and its stream.
 *   - plain-pointer calls are synchronous on return.  Evaluations carry up to 8 parameter vectors
 *     in the kernel arguments and receive the statistics through mapped pinned host memory that the
 *     kernel's last block writes and the calling thread polls: no H2D / D2H copy, no stream sync.
 *     *_device calls take device pointers and a cudaStream_t (as void*), are asynchronous, and are
 *     what an on-device sweep/MCMC loop uses.
 *   - a handle is not thread-safe; distinct handles are independent.
 *   - device memory comes from the device's default stream-ordered pool (cudaMallocAsync on the handle's stream; the
 *     pool's release threshold is raised to 8 GB at nngp_create), so the blocks of a destroyed handle -- and the
 *     pinned result lines, and the per-device exp table -- serve the next one: a model can be rebuilt without paying
 *     for cudaMalloc / cudaFree again.  If a *_device entry point ran on a caller's stream, the next call that
 *     releases or replaces buffers synchronises the whole device first.
 *   - there is NO CPU fallback: without a CUDA device nngp_create fails with NNGP_ENODEVICE.
 */
#ifndef NNGP_B200_H
#define NNGP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nngp_handle nngp_handle;

enum {
    NNGP_OK = 0,
    NNGP_EINVAL = 1,    /* bad argument (NULL, m out of range, D unsupported, ...) */
    NNGP_ENODEVICE = 2, /* no CUDA device / wrong architecture */
    NNGP_ECUDA = 3,     /* a CUDA runtime call failed; see nngp_last_error */
    NNGP_ESTATE = 4     /* call order: data / neighbours not set yet */
};

enum { NNGP_F64 = 0, NNGP_F32 = 1 };                               /* arithmetic of stages 2-3 */
enum { NNGP_EXPONENTIAL = 0, NNGP_MATERN32 = 1, NNGP_MATERN52 = 2 }; /* kernel_id */

#define NNGP_MAX_M 32   /* neighbours per location (north_star: m <= 32) */
#define NNGP_MAX_D 3    /* spatial dimension of the coordinates */
#define NNGP_NPARAM 4   /* one parameter vector = {sigma2, phi, tau2, nu(reserved)} */
#define NNGP_NSTAT 3    /* one result = {sum log F_i, sum r_i^2/F_i, n_bad} */

/* Library / device -------------------------------------------------------------------------- */
const char *nngp_version(void);
const char *nngp_last_error(const nngp_handle *h);

/* Creates an engine on CUDA device `device` computing stages 2-3 in `dtype` (NNGP_F64|NNGP_F32;
 * stage 1 is always fp64).  Replaces: object construction, nngp.py:6-12. */
int nngp_create(nngp_handle **h, int device, int dtype);
/* One handle over `ndev` (1..8) distinct devices of this process (nngp_create(h, devices, ndev,
 * dtype)).  The devices must have P2P access to each other (one NVLink / NVSwitch domain).  Every entry point
 * below accepts it unless it says otherwise: data is replicated, the shard and the neighbour table are split in
 * contiguous blocks over the devices, evaluations return the total.  Device-pointer entry points
 * (nngp_loglik_device*, nngp_peer_*, nngp_neighbors_device_ptr) and nngp_build_neighbors /
 * nngp_build_neighbors_capped are for device handles only (NNGP_ESTATE). */
int nngp_create_multi(nngp_handle **h, const int *devices, int ndev, int dtype);
int nngp_device_count(const nngp_handle *h);  /* devices behind this handle */
int nngp_visible_devices(void);               /* sm_100 devices this process can see (0 without a driver) */
void nngp_destroy(nngp_handle *h);

/* Uploads the reference set and the response.  coords: n x D row-major fp64 (the reference's
 * `s`, nngp.py:29-31 'S=T' branch: s = t); y: n fp64 (nngp.py:8); eps2: n fp64 per-observation
 * variances added to the diagonal (the square of nngp.py:9's `eps`), or NULL.
 * Resets the shard to [0, n) and drops any neighbour table. */
int nngp_set_data(nngp_handle *h, const double *coords, int64_t n, int D, const double *y,
                  const double *eps2);
/* Replaces y only (n fp64), keeping coordinates and neighbours. */
int nngp_set_y(nngp_handle *h, const double *y);
/* Replaces the per-observation variances only (n fp64, the square of nngp.py:9's `eps`), keeping coordinates,
 * y and neighbours.  A latent-field density (reference sites carry no nugget, observations do) changes these
 * with tau2 while everything else stays resident. */
int nngp_set_eps2(nngp_handle *h, const double *eps2);

/* Rows [lo, hi) of the ordering are this handle's shard: stages 2-3 and the returned partial
 * statistics cover exactly these rows.  Coordinates / y stay replicated in full. */
int nngp_set_shard(nngp_handle *h, int64_t lo, int64_t hi);

/* Stage 1 -- ordered k-NN.  Replaces _make_s_neighbor_sets, nngp.py:49-62: for each i the
 * min(m, i) nearest j < i in ascending (d2, j), d2 the sklearn fp64 squared distance; row i of the
 * device table (n x m int32, -1 padded) is filled.  Query tiles (NNGP_KNN_TILE consecutive rows)
 * are taken heaviest first; with tile_stride > 1 only tiles whose rank from the heavy end is
 * congruent to tile_offset are computed (balanced multi-GPU split) and every other row is set to
 * NNGP_ROW_UNSET so an elementwise MAX across ranks assembles the table. */
#define NNGP_KNN_TILE 128
#define NNGP_ROW_UNSET (-2)
int nngp_build_neighbors(nngp_handle *h, int m, int tile_offset, int tile_stride);
/* Stage 1 in sub-quadratic time (same contract and bit-identical output: _make_s_neighbor_sets,
 * nngp.py:49-62): rows [row_lo, row_hi) of the table are searched over a uniform cell grid, level by
 * level of the ordering (knn_grid.cu); rows outside the range hold NNGP_ROW_UNSET or their correct
 * value, so ranks that each build their own shard assemble the table with an elementwise MAX.
 * algo: NNGP_KNN_AUTO = grid unless the coordinates are non-finite or so clustered that the cell
 * histogram predicts more work than brute force; NNGP_KNN_GRID / NNGP_KNN_BRUTE force one. */
enum { NNGP_KNN_AUTO = 0, NNGP_KNN_GRID = 1, NNGP_KNN_BRUTE = 2 };
int nngp_build_neighbors_grid(nngp_handle *h, int m, int64_t row_lo, int64_t row_hi, int algo);
/* The same search for the rows of the handle's shard [lo, hi) ONLY, and the table keeps only those rows
 * ((hi - lo) x m instead of n x m: what a rank of a multi-GPU run needs -- SURVEY 8 e2).  Evaluations,
 * nngp_factors, nngp_cov_blocks and nngp_get_neighbor_rows then work on rows inside that window;
 * nngp_get_neighbors needs a full table.  nngp_neighbor_window reports the rows held. */
int nngp_build_neighbors_shard(nngp_handle *h, int m, int algo);
int nngp_neighbor_window(const nngp_handle *h, int64_t *row0, int64_t *rows);
/* Same search with the candidates of row i restricted to j < min(i, cand_cap).  Replaces
 * _make_t_neighbor_sets for S != T, nngp.py:68-71 (KDTree(s).query(t_i, m): the sites of t appended after
 * the nRef rows of s, row_lo = cand_cap = nRef).  With the n reference
 * sites in rows [0, n) and q prediction sites appended as rows [n, n + q), row_lo = cand_cap = n gives
 * every new site its m nearest REFERENCE sites (the neighbour sets kriging conditions on; the step after
 * the path); nngp_factors over the same rows then returns the kriging weights and variances. */
int nngp_build_neighbors_capped(nngp_handle *h, int m, int64_t row_lo, int64_t row_hi, int64_t cand_cap, int algo);
/* Grid search knobs: cells hold lambda_scale * (m + 2 sqrt m) / unit-ball-volume usable points
 * (default 0.5); rows below brute_rows (default 128) always use brute force. */
int nngp_set_knn_tuning(nngp_handle *h, double lambda_scale, int64_t brute_rows);
/* 1 if the last nngp_build_neighbors_grid call went through the grid, 0 if it fell back. */
int nngp_knn_used_grid(const nngp_handle *h);
/* Injects a table (n x m int32 row-major, -1 padded; valid entries first in each row).  The table is checked on
 * the device before it is accepted: every entry of row i must lie in [-1, i) -- neighbours precede their row, so no
 * gather can leave the records -- with the padding at the tail; otherwise NNGP_EINVAL. */
int nngp_set_neighbors(nngp_handle *h, const int32_t *idx, int m);
int nngp_get_neighbors(nngp_handle *h, int32_t *out);
/* Rows [i0, i1) of the table only: out is (i1-i0) x m int32. */
int nngp_get_neighbor_rows(nngp_handle *h, int64_t i0, int64_t i1, int32_t *out);
/* Plain k-NN over all rows, the query itself included (no ordering constraint), same (d2, j) order;
 * out: n x k int32 on the host.  Replaces the neighbour search inside _init_ws, nngp.py:45-47
 * (KNeighborsRegressor(5).fit(t, y).predict(s): the mean of y over these rows is `ws`). */
int nngp_knn_plain(nngp_handle *h, int k, int32_t *out);
/* Device address of the n x m int32 table (for an NCCL exchange by the caller), or NULL when the handle holds
 * a window only; nngp_neighbor_window_device_ptr is the address of the first row HELD (see nngp_neighbor_window). */
void *nngp_neighbors_device_ptr(nngp_handle *h);
void *nngp_neighbor_window_device_ptr(nngp_handle *h);

/* Stages 2-3 fused -- the metric's call.  For each of K parameter vectors
 * (params: K x NNGP_NPARAM fp64) builds C_N(i), c_i, C(i,i) (_CNs nngp.py:78-82, _Ccross
 * nngp.py:84-86, _Cs nngp.py:92-96), factorises (_Bsi nngp.py:73-76, _Fsi nngp.py:88-90) and
 * reduces over the shard.  out: K x NNGP_NSTAT fp64 = {sum log F_i, sum (y_i - b_i^T y_N(i))^2/F_i,
 * n_bad}; locations whose factorisation met a non-positive pivot are counted in n_bad and left out
 * of the sums. */
int nngp_loglik(nngp_handle *h, int kernel_id, const double *params, int K, double *out);
/* One parameter vector by value: out3 = {sum log F_i, sum r_i^2/F_i, n_bad}.  The shortest host path of the
 * metric's call (an MCMC step): no staging buffers, no copies.  The result is the total over all GPUs when the
 * handle is a multi-device handle or its peer exchange is connected (every rank must then call it), otherwise
 * the handle's shard. */
int nngp_loglik_terms(nngp_handle *h, int kernel_id, double sigma2, double phi, double tau2, double *out3);
/* Same with device pointers on `stream` (NULL = the handle's stream); no host synchronisation. */
int nngp_loglik_device(nngp_handle *h, int kernel_id, const double *d_params, int K,
                       double *d_out, void *stream);

/* Multi-GPU evaluation with the sum over ranks fused into the kernel (one process per GPU of ONE node).
 * Each rank calls nngp_peer_export (allocates its exchange buffer for up to K_cap parameter vectors and
 * returns its CUDA IPC handle, NNGP_IPC_HANDLE_BYTES bytes), the caller gathers the `world` handles in
 * rank order (any transport; pynngp_b200 uses torch.distributed), and every rank calls nngp_peer_connect.
 * nngp_loglik_device_allreduce then behaves like nngp_loglik_device except that d_out receives the sum over
 * ALL ranks, bitwise identical on every rank: the last block of each rank stores its K x 3 partials into
 * every peer's buffer over NVLink (P2P stores of self-stamped 16-byte lines: no flag, no fence), waits for all
 * ranks' lines and sums in rank order -- no second launch, no NCCL call.  Every rank must issue the same
 * sequence of these calls.  If a peer does not arrive within ~30 s the statistics come back as NaN.
 * K_cap must not exceed the device's SM count (one waiting block per parameter vector must stay resident). */
#define NNGP_IPC_HANDLE_BYTES 64
int nngp_peer_export(nngp_handle *h, int K_cap, unsigned char *handle_out);
int nngp_peer_connect(nngp_handle *h, int rank, int world, const unsigned char *handles);
int nngp_loglik_device_allreduce(nngp_handle *h, int kernel_id, const double *d_params, int K,
                                 double *d_out, void *stream);
/* Host-pointer form (synchronous, like nngp_loglik): out receives the K x 3 totals over all ranks. */
int nngp_loglik_allreduce(nngp_handle *h, int kernel_id, const double *params, int K, double *out);

/* Per-location factors for rows [i0, i1) (any rows, not only the shard): B (i1-i0) x m fp64
 * (b_i = C_N(i)^-1 c_i, zero padded; _Bsi nngp.py:73-76), F (i1-i0) (_Fsi nngp.py:88-90).
 * Either output may be NULL. */
int nngp_factors(nngp_handle *h, int kernel_id, const double *params, int64_t i0, int64_t i1,
                 double *B, double *F);
/* Covariance blocks for rows [i0, i1): CN (i1-i0) x m x m (zero outside the leading p x p; _CNs
 * nngp.py:78-82), cc (i1-i0) x m (_Ccross nngp.py:84-86), cs (i1-i0) (_Cs nngp.py:92-96).  Any
 * output may be NULL. */
int nngp_cov_blocks(nngp_handle *h, int kernel_id, const double *params, int64_t i0, int64_t i1,
                    double *CN, double *cc, double *cs);

/* Introspection for the bench: number of kernels this handle has launched so far, and a measured
 * FP64 (dtype NNGP_F64) or FP32 FMA peak of the device in thread-instructions per second
 * (register-resident dependent-chain microbenchmark, `iters` FMAs per thread). */
int64_t nngp_launch_count(const nngp_handle *h);
/* Device-side timing of the evaluation kernels for a bench: with timing on, every evaluation launch is
 * bracketed by CUDA events on its stream; nngp_last_eval_ms waits for the last one and returns its duration
 * (for a multi-device handle the maximum over the devices -- each kernel's tail includes the wait for its
 * peers, so this is the device time of the whole exchange-terminated step). */
int nngp_set_timing(nngp_handle *h, int on);
int nngp_last_eval_ms(nngp_handle *h, double *ms);
int nngp_measure_fma_peak(nngp_handle *h, int dtype, int iters, double *instr_per_s);

#ifdef __cplusplus
}
#endif
#endif /* NNGP_B200_H */
