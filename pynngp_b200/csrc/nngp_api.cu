// nngp_api.cu -- the C ABI of libnngp_b200.so (include/nngp_b200.h): handle, uploads, dispatch.
// No CPU fallback exists anywhere in this library: every compute entry point launches a kernel.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <cmath>
#include <vector>

#include "nngp_common.cuh"

// one launcher/occupancy pair per (dtype, correlation family) translation unit
#define NNGP_DECLARE_FAMILY(NAME)                                                                \
    cudaError_t nngp_launch_##NAME(int m, int D, const EvalArgs &a, int K, int grid_x,           \
                                   cudaStream_t stream);                                         \
    void nngp_shape_##NAME(int m, int D, int *blocks_per_sm, int *loc_per_warp);
NNGP_DECLARE_FAMILY(f64_exp)
NNGP_DECLARE_FAMILY(f64_m32)
NNGP_DECLARE_FAMILY(f64_m52)
NNGP_DECLARE_FAMILY(f32_exp)
NNGP_DECLARE_FAMILY(f32_m32)
NNGP_DECLARE_FAMILY(f32_m52)

namespace {

std::string g_create_error;

int fail(nngp_handle *h, int code, const std::string &msg)
{
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

int cuda_fail(nngp_handle *h, cudaError_t e, const char *what)
{
    return fail(h, NNGP_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CUDA_TRY(h, call)                                                  \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess) return cuda_fail(h, e_, #call);             \
    } while (0)

template <typename P>
void free_dev(P *&p)
{
    if (p) cudaFree(p);
    p = nullptr;
}

void family_shape(int dtype, int kernel_id, int m, int D, int *per_sm, int *lpw)
{
    if (dtype == NNGP_F64) {
        if (kernel_id == NNGP_EXPONENTIAL) return nngp_shape_f64_exp(m, D, per_sm, lpw);
        if (kernel_id == NNGP_MATERN32) return nngp_shape_f64_m32(m, D, per_sm, lpw);
        return nngp_shape_f64_m52(m, D, per_sm, lpw);
    }
    if (kernel_id == NNGP_EXPONENTIAL) return nngp_shape_f32_exp(m, D, per_sm, lpw);
    if (kernel_id == NNGP_MATERN32) return nngp_shape_f32_m32(m, D, per_sm, lpw);
    return nngp_shape_f32_m52(m, D, per_sm, lpw);
}

cudaError_t family_launch(int dtype, int kernel_id, int m, int D, const EvalArgs &a, int K, int grid_x,
                          cudaStream_t stream)
{
    if (dtype == NNGP_F64) {
        if (kernel_id == NNGP_EXPONENTIAL) return nngp_launch_f64_exp(m, D, a, K, grid_x, stream);
        if (kernel_id == NNGP_MATERN32) return nngp_launch_f64_m32(m, D, a, K, grid_x, stream);
        return nngp_launch_f64_m52(m, D, a, K, grid_x, stream);
    }
    if (kernel_id == NNGP_EXPONENTIAL) return nngp_launch_f32_exp(m, D, a, K, grid_x, stream);
    if (kernel_id == NNGP_MATERN32) return nngp_launch_f32_m32(m, D, a, K, grid_x, stream);
    return nngp_launch_f32_m52(m, D, a, K, grid_x, stream);
}

// grid.x for nloc locations: enough resident blocks to fill every SM, never more than the work.  The
// occupancy query is cached per (kernel family, m, D): it costs several microseconds of host time.
int grid_for(nngp_handle *h, int kernel_id, int64_t nloc)
{
    const int key = ((kernel_id * 64 + h->m) * 4 + h->D) * 2 + h->dtype;
    if (h->shape_key != key) {
        int per_sm = 1, lpw = 8;
        family_shape(h->dtype, kernel_id, h->m, h->D, &per_sm, &lpw);
        h->shape_key = key; h->shape_per_sm = per_sm; h->shape_lpw = lpw;
    }
    const int64_t groups = (nloc + h->shape_lpw - 1) / h->shape_lpw;
    const int64_t need = (groups + 3) / 4;  // 4 warps per block
    int64_t g = int64_t(h->num_sms) * h->shape_per_sm;
    if (g > need) g = need;
    if (g < 1) g = 1;
    return int(g);
}

int ensure_scratch(nngp_handle *h, int K, int grid)
{
    if (K > h->K_cap) {
        free_dev(h->d_params); free_dev(h->d_out); free_dev(h->d_counters);
        if (h->h_stage) { cudaFreeHost(h->h_stage); h->h_stage = nullptr; }
        int cap = K < 16 ? 16 : K;
        CUDA_TRY(h, cudaMalloc(&h->d_params, sizeof(double) * NNGP_NPARAM * cap));
        CUDA_TRY(h, cudaMalloc(&h->d_out, sizeof(double) * NNGP_NSTAT * cap));
        CUDA_TRY(h, cudaMalloc(&h->d_counters, sizeof(unsigned int) * cap));
        CUDA_TRY(h, cudaMemset(h->d_counters, 0, sizeof(unsigned int) * cap));
        CUDA_TRY(h, cudaMallocHost(&h->h_stage, sizeof(double) * (NNGP_NPARAM + NNGP_NSTAT) * cap));
        free_dev(h->d_partials);
        h->grid_cap = 0;
        h->K_cap = cap;
    }
    if (grid > h->grid_cap || !h->d_partials) {
        free_dev(h->d_partials);
        int gc = grid < 1024 ? 1024 : grid;
        CUDA_TRY(h, cudaMalloc(&h->d_partials, sizeof(double) * 3 * size_t(gc) * h->K_cap));
        h->grid_cap = gc;
    }
    return NNGP_OK;
}

int check_eval(nngp_handle *h, int kernel_id, const void *params, int K)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!h->pts) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
    if (!h->has_nbr) return fail(h, NNGP_ESTATE, "no neighbour table: call nngp_build_neighbors or nngp_set_neighbors");
    if (kernel_id < 0 || kernel_id > NNGP_MATERN52) return fail(h, NNGP_EINVAL, "unknown kernel_id");
    if (!params || K < 1) return fail(h, NNGP_EINVAL, "params must hold K >= 1 parameter vectors");
    return NNGP_OK;
}

}  // namespace

extern "C" {

const char *nngp_version(void) { return "nngp_b200 0.1 (sm_100a)"; }

const char *nngp_last_error(const nngp_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int nngp_create(nngp_handle **out, int device, int dtype)
{
    if (!out) return fail(nullptr, NNGP_EINVAL, "null handle pointer");
    *out = nullptr;
    if (dtype != NNGP_F64 && dtype != NNGP_F32) return fail(nullptr, NNGP_EINVAL, "dtype must be NNGP_F64 or NNGP_F32");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, NNGP_ENODEVICE,
                    std::string("no CUDA device (this library has no CPU fallback): ") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= ndev) return fail(nullptr, NNGP_EINVAL, "device index out of range");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return fail(nullptr, NNGP_ENODEVICE, "libnngp_b200 is built for sm_100a only; device is sm_" +
                                                 std::to_string(prop.major) + std::to_string(prop.minor));
    if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    nngp_handle *h = new nngp_handle();
    h->device = device;
    h->dtype = dtype;
    h->num_sms = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete h;
        return cuda_fail(nullptr, e, "cudaStreamCreate");
    }
    if ((e = cudaMalloc(&h->d_tile_counter, sizeof(unsigned int))) != cudaSuccess) {
        cudaStreamDestroy(h->stream);
        delete h;
        return cuda_fail(nullptr, e, "cudaMalloc");
    }
    {   // parameter-independent part of the covariance build's exp table: 2^(j/2048), j < 2048
        std::vector<double> tab(2048);
        for (int j = 0; j < 2048; ++j) tab[j] = exp2(double(j) / 2048.0);
        if ((e = cudaMalloc(&h->d_exp2tab, sizeof(double) * 2048)) != cudaSuccess ||
            (e = cudaMemcpy(h->d_exp2tab, tab.data(), sizeof(double) * 2048, cudaMemcpyHostToDevice)) != cudaSuccess) {
            cudaFree(h->d_tile_counter);
            cudaStreamDestroy(h->stream);
            delete h;
            return cuda_fail(nullptr, e, "cudaMalloc");
        }
    }
    *out = h;
    return NNGP_OK;
}

void nngp_destroy(nngp_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    free_dev(h->pts); free_dev(h->eps2); free_dev(h->nbr); free_dev(h->d_ystage);
    free_dev(h->d_params); free_dev(h->d_out); free_dev(h->d_partials);
    free_dev(h->d_counters); free_dev(h->d_tile_counter); free_dev(h->d_exp2tab);
    for (int r = 0; r < NNGP_MAX_PEERS; ++r)
        if (h->peer_base[r] && h->peer_base[r] != h->xbuf) cudaIpcCloseMemHandle(h->peer_base[r]);
    free_dev(h->xbuf);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    cudaStreamDestroy(h->stream);
    delete h;
}

int nngp_set_data(nngp_handle *h, const double *coords, int64_t n, int D, const double *y, const double *eps2)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!coords || !y) return fail(h, NNGP_EINVAL, "coords and y must not be NULL");
    if (n < 1 || n > 2147483647LL) return fail(h, NNGP_EINVAL, "n must be in [1, 2^31)");
    if (D < 1 || D > NNGP_MAX_D) return fail(h, NNGP_EINVAL, "D must be 1, 2 or 3");
    CUDA_TRY(h, cudaSetDevice(h->device));
    free_dev(h->pts); free_dev(h->eps2); free_dev(h->nbr); free_dev(h->d_ystage);  // the landing buffer is sized by n
    h->has_nbr = false; h->m = 0;
    h->n = n; h->D = D; h->lo = 0; h->hi = n;
    // raw arrays up, packed into {x, y, z, yval} records on the device (pack.cu), which also reduces the
    // bounding box for the grid search of stage 1
    double *d_coords = nullptr, *d_y = nullptr, *d_e2 = nullptr;
    auto cleanup = [&]() { free_dev(d_coords); free_dev(d_y); if (d_e2 != h->eps2) free_dev(d_e2); };
#define SET_TRY(call)                                                          \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) { cleanup(); free_dev(h->eps2); free_dev(h->pts); return cuda_fail(h, e_, #call); }  \
    } while (0)
    SET_TRY(cudaMalloc(&h->pts, sizeof(double4) * (size_t)n));
    SET_TRY(cudaMalloc(&d_coords, sizeof(double) * (size_t)n * D));
    SET_TRY(cudaMalloc(&d_y, sizeof(double) * (size_t)n));
    SET_TRY(cudaMemcpyAsync(d_coords, coords, sizeof(double) * (size_t)n * D, cudaMemcpyHostToDevice, h->stream));
    SET_TRY(cudaMemcpyAsync(d_y, y, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    if (eps2) {
        SET_TRY(cudaMalloc(&d_e2, sizeof(double) * (size_t)n));
        SET_TRY(cudaMemcpyAsync(d_e2, eps2, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
        if (D > 2) h->eps2 = d_e2;  // D = 3: kept as its own array; D < 3: folded into the records' z slot
    }
    SET_TRY(launch_pack_records(h, d_coords, d_y, d_e2, h->stream));  // synchronises the stream
#undef SET_TRY
    cleanup();
    return NNGP_OK;
}

int nngp_set_y(nngp_handle *h, const double *y)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!h->pts) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
    if (!y) return fail(h, NNGP_EINVAL, "y must not be NULL");
    CUDA_TRY(h, cudaSetDevice(h->device));
    // one contiguous copy, then a kernel writes the yval lane of the records (pack.cu)
    if (!h->d_ystage) CUDA_TRY(h, cudaMalloc(&h->d_ystage, sizeof(double) * (size_t)h->n));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_ystage, y, sizeof(double) * (size_t)h->n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, launch_scatter_lane(h, h->d_ystage, 3, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NNGP_OK;
}

int nngp_set_eps2(nngp_handle *h, const double *eps2)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!h->pts) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
    if (!eps2) return fail(h, NNGP_EINVAL, "eps2 must not be NULL");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->D > 2) {  // D = 3: its own array (the records' z slot holds a coordinate)
        if (!h->eps2) CUDA_TRY(h, cudaMalloc(&h->eps2, sizeof(double) * (size_t)h->n));
        CUDA_TRY(h, cudaMemcpyAsync(h->eps2, eps2, sizeof(double) * (size_t)h->n, cudaMemcpyHostToDevice, h->stream));
    } else {         // D < 3: the records' z slot
        if (!h->d_ystage) CUDA_TRY(h, cudaMalloc(&h->d_ystage, sizeof(double) * (size_t)h->n));
        CUDA_TRY(h, cudaMemcpyAsync(h->d_ystage, eps2, sizeof(double) * (size_t)h->n, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, launch_scatter_lane(h, h->d_ystage, 2, h->stream));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NNGP_OK;
}

int nngp_set_shard(nngp_handle *h, int64_t lo, int64_t hi)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!h->pts) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
    if (lo < 0 || hi < lo || hi > h->n) return fail(h, NNGP_EINVAL, "shard must satisfy 0 <= lo <= hi <= n");
    h->lo = lo; h->hi = hi;
    return NNGP_OK;
}

static int alloc_nbr(nngp_handle *h, int m)
{
    if (m < 1 || m > NNGP_MAX_M) return fail(h, NNGP_EINVAL, "m must be in [1, 32]");
    if (h->nbr && h->m != m) free_dev(h->nbr);
    if (!h->nbr) CUDA_TRY(h, cudaMalloc(&h->nbr, sizeof(int32_t) * (size_t)h->n * m));
    h->m = m;
    return NNGP_OK;
}

int nngp_build_neighbors(nngp_handle *h, int m, int tile_offset, int tile_stride)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!h->pts) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
    if (tile_stride < 1 || tile_offset < 0 || tile_offset >= tile_stride)
        return fail(h, NNGP_EINVAL, "need 0 <= tile_offset < tile_stride");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = alloc_nbr(h, m);
    if (rc) return rc;
    CUDA_TRY(h, launch_knn_ordered(h, m, tile_offset, tile_stride, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->has_nbr = true;
    return NNGP_OK;
}

static int build_neighbors_capped(nngp_handle *h, int m, int64_t row_lo, int64_t row_hi, int64_t cand_cap, int algo)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!h->pts) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
    if (row_lo < 0 || row_hi < row_lo || row_hi > h->n) return fail(h, NNGP_EINVAL, "need 0 <= row_lo <= row_hi <= n");
    if (cand_cap < 1) return fail(h, NNGP_EINVAL, "cand_cap must be >= 1");
    if (algo != NNGP_KNN_AUTO && algo != NNGP_KNN_GRID && algo != NNGP_KNN_BRUTE)
        return fail(h, NNGP_EINVAL, "algo must be NNGP_KNN_AUTO, NNGP_KNN_GRID or NNGP_KNN_BRUTE");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = alloc_nbr(h, m);
    if (rc) return rc;
    int used = 0;
    if (algo != NNGP_KNN_BRUTE) {
        CUDA_TRY(h, launch_knn_grid(h, true, m, row_lo, row_hi, cand_cap, h->nbr, h->stream, algo == NNGP_KNN_GRID, &used));
        if (!used && algo == NNGP_KNN_GRID) return fail(h, NNGP_EINVAL, "grid search needs finite coordinates");
    }
    if (!used) {
        // brute force over the query tiles covering [row_lo, row_hi); every other row is unset
        const int64_t t_lo = row_lo / NNGP_KNN_TILE * NNGP_KNN_TILE;
        CUDA_TRY(h, launch_fill_i32(h, h->nbr, t_lo * m, NNGP_ROW_UNSET, h->stream));
        CUDA_TRY(h, launch_knn_brute_rows(h, m, t_lo, row_hi, cand_cap, h->nbr, h->stream));
        CUDA_TRY(h, launch_fill_i32(h, h->nbr + row_hi * m, (h->n - row_hi) * m, NNGP_ROW_UNSET, h->stream));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->knn_used_grid = used;
    h->has_nbr = true;
    return NNGP_OK;
}

int nngp_build_neighbors_grid(nngp_handle *h, int m, int64_t row_lo, int64_t row_hi, int algo)
{
    return build_neighbors_capped(h, m, row_lo, row_hi, INT64_MAX, algo);
}

int nngp_build_neighbors_capped(nngp_handle *h, int m, int64_t row_lo, int64_t row_hi, int64_t cand_cap, int algo)
{
    return build_neighbors_capped(h, m, row_lo, row_hi, cand_cap, algo);
}

int nngp_set_knn_tuning(nngp_handle *h, double lambda_scale, int64_t brute_rows)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!(lambda_scale > 0.0) || brute_rows < 1) return fail(h, NNGP_EINVAL, "need lambda_scale > 0 and brute_rows >= 1");
    h->knn_lambda_scale = lambda_scale;
    h->knn_brute_rows = brute_rows;
    return NNGP_OK;
}

int nngp_knn_used_grid(const nngp_handle *h) { return h ? h->knn_used_grid : 0; }

int nngp_set_neighbors(nngp_handle *h, const int32_t *idx, int m)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!h->pts) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
    if (!idx) return fail(h, NNGP_EINVAL, "idx must not be NULL");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = alloc_nbr(h, m);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(h->nbr, idx, sizeof(int32_t) * (size_t)h->n * m, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->has_nbr = true;
    return NNGP_OK;
}

int nngp_get_neighbors(nngp_handle *h, int32_t *out)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!h->has_nbr) return fail(h, NNGP_ESTATE, "no neighbour table");
    if (!out) return fail(h, NNGP_EINVAL, "out must not be NULL");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaMemcpyAsync(out, h->nbr, sizeof(int32_t) * (size_t)h->n * h->m, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NNGP_OK;
}

int nngp_get_neighbor_rows(nngp_handle *h, int64_t i0, int64_t i1, int32_t *out)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!h->has_nbr) return fail(h, NNGP_ESTATE, "no neighbour table");
    if (!out || i0 < 0 || i1 < i0 || i1 > h->n) return fail(h, NNGP_EINVAL, "need out != NULL and 0 <= i0 <= i1 <= n");
    if (i1 == i0) return NNGP_OK;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaMemcpyAsync(out, h->nbr + i0 * h->m, sizeof(int32_t) * size_t(i1 - i0) * h->m, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NNGP_OK;
}

int nngp_knn_plain(nngp_handle *h, int k, int32_t *out)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!h->pts) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
    if (!out || k < 1 || k > NNGP_MAX_M) return fail(h, NNGP_EINVAL, "need out != NULL and 1 <= k <= 32");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int32_t *d_tab = nullptr;
    CUDA_TRY(h, cudaMalloc(&d_tab, sizeof(int32_t) * (size_t)h->n * k));
    int used = 0;
    cudaError_t e = launch_knn_grid(h, false, k, 0, h->n, INT64_MAX, d_tab, h->stream, 0, &used);
    if (e == cudaSuccess && !used) e = launch_knn_plain(h, k, d_tab, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_tab, sizeof(int32_t) * (size_t)h->n * k, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_tab);
    if (e != cudaSuccess) return cuda_fail(h, e, "nngp_knn_plain");
    return NNGP_OK;
}

void *nngp_neighbors_device_ptr(nngp_handle *h) { return (h && h->has_nbr) ? (void *)h->nbr : nullptr; }

int nngp_loglik_device(nngp_handle *h, int kernel_id, const double *d_params, int K, double *d_out, void *stream)
{
    int rc = check_eval(h, kernel_id, d_params, K);
    if (rc) return rc;
    if (!d_out) return fail(h, NNGP_EINVAL, "d_out must not be NULL");
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    const int64_t nloc = h->hi - h->lo;
    if (nloc == 0) {  // empty shard: the statistics are exactly zero
        CUDA_TRY(h, cudaMemsetAsync(d_out, 0, sizeof(double) * NNGP_NSTAT * K, st));
        return NNGP_OK;
    }
    const int grid = grid_for(h, kernel_id, nloc);
    if ((rc = ensure_scratch(h, K, grid))) return rc;
    EvalArgs a{};
    a.pts = h->pts; a.eps2 = h->eps2; a.nbr = h->nbr;
    a.lo = h->lo; a.hi = h->hi; a.m = h->m;
    a.params = d_params; a.partials = h->d_partials; a.counters = h->d_counters; a.out = d_out;
    a.emit = 0; a.exp2tab = h->d_exp2tab; a.K = K;
    CUDA_TRY(h, family_launch(h->dtype, kernel_id, h->m, h->D, a, K, grid, st));
    ++h->launches;
    return NNGP_OK;
}

static size_t peer_slots_bytes(int K_cap) { return sizeof(double) * 2 * NNGP_MAX_PEERS * size_t(K_cap) * 3; }
static size_t peer_flags_bytes(int K_cap) { return sizeof(unsigned long long) * 2 * NNGP_MAX_PEERS * size_t(K_cap); }

int nngp_peer_export(nngp_handle *h, int K_cap, unsigned char *handle_out)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!handle_out || K_cap < 1) return fail(h, NNGP_EINVAL, "need handle_out != NULL and K_cap >= 1");
    if (h->px.world > 1) return fail(h, NNGP_ESTATE, "peer exchange is already connected");
    CUDA_TRY(h, cudaSetDevice(h->device));
    free_dev(h->xbuf);
    const size_t bytes = peer_slots_bytes(K_cap) + peer_flags_bytes(K_cap);
    CUDA_TRY(h, cudaMalloc(&h->xbuf, bytes));
    CUDA_TRY(h, cudaMemset(h->xbuf, 0, bytes));
    CUDA_TRY(h, cudaDeviceSynchronize());
    h->px.K_cap = K_cap;
    cudaIpcMemHandle_t ih;
    CUDA_TRY(h, cudaIpcGetMemHandle(&ih, h->xbuf));
    static_assert(sizeof(ih) == NNGP_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    memcpy(handle_out, &ih, sizeof(ih));
    return NNGP_OK;
}

int nngp_peer_connect(nngp_handle *h, int rank, int world, const unsigned char *handles)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!h->xbuf) return fail(h, NNGP_ESTATE, "call nngp_peer_export first");
    if (!handles || world < 2 || world > NNGP_MAX_PEERS || rank < 0 || rank >= world)
        return fail(h, NNGP_EINVAL, "need 2 <= world <= 8, 0 <= rank < world and world handles");
    CUDA_TRY(h, cudaSetDevice(h->device));
    for (int r = 0; r < world; ++r) {
        void *base = h->xbuf;
        if (r != rank) {
            cudaIpcMemHandle_t ih;
            memcpy(&ih, handles + size_t(r) * NNGP_IPC_HANDLE_BYTES, sizeof(ih));
            cudaError_t e = cudaIpcOpenMemHandle(&base, ih, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                for (int q = 0; q < r; ++q)
                    if (h->peer_base[q] && h->peer_base[q] != h->xbuf) { cudaIpcCloseMemHandle(h->peer_base[q]); h->peer_base[q] = nullptr; }
                return cuda_fail(h, e, "cudaIpcOpenMemHandle (peers must be GPUs of one node with P2P access)");
            }
        }
        h->peer_base[r] = base;
        h->px.slots[r] = reinterpret_cast<double *>(base);
        h->px.flags[r] = reinterpret_cast<unsigned long long *>(static_cast<char *>(base) + peer_slots_bytes(h->px.K_cap));
    }
    h->px.rank = rank;
    h->px.world = world;
    h->px.gen = 0;
    return NNGP_OK;
}

int nngp_loglik_device_allreduce(nngp_handle *h, int kernel_id, const double *d_params, int K, double *d_out, void *stream)
{
    int rc = check_eval(h, kernel_id, d_params, K);
    if (rc) return rc;
    if (!d_out) return fail(h, NNGP_EINVAL, "d_out must not be NULL");
    if (h->px.world < 2) return fail(h, NNGP_ESTATE, "peer exchange is not connected (nngp_peer_export / nngp_peer_connect)");
    if (K > h->px.K_cap) return fail(h, NNGP_EINVAL, "K exceeds the exchange buffer's K_cap");
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    PeerExchange px = h->px;
    px.gen = ++h->px.gen;  // every rank issues the same sequence of exchanges
    const int64_t nloc = h->hi - h->lo;
    if (nloc == 0) {
        CUDA_TRY(h, launch_peer_zero(h, px, K, d_out, st));
        ++h->launches;
        return NNGP_OK;
    }
    const int grid = grid_for(h, kernel_id, nloc);
    if ((rc = ensure_scratch(h, K, grid))) return rc;
    EvalArgs a{};
    a.pts = h->pts; a.eps2 = h->eps2; a.nbr = h->nbr;
    a.lo = h->lo; a.hi = h->hi; a.m = h->m;
    a.params = d_params; a.partials = h->d_partials; a.counters = h->d_counters; a.out = d_out;
    a.emit = 0; a.exp2tab = h->d_exp2tab; a.K = K;
    a.px = px;
    CUDA_TRY(h, family_launch(h->dtype, kernel_id, h->m, h->D, a, K, grid, st));
    ++h->launches;
    return NNGP_OK;
}

static int loglik_host(nngp_handle *h, int kernel_id, const double *params, int K, double *out, bool allreduce)
{
    int rc = check_eval(h, kernel_id, params, K);
    if (rc) return rc;
    if (!out) return fail(h, NNGP_EINVAL, "out must not be NULL");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if ((rc = ensure_scratch(h, K, 1))) return rc;
    double *hp = h->h_stage, *ho = h->h_stage + size_t(NNGP_NPARAM) * h->K_cap;
    memcpy(hp, params, sizeof(double) * NNGP_NPARAM * K);
    CUDA_TRY(h, cudaMemcpyAsync(h->d_params, hp, sizeof(double) * NNGP_NPARAM * K, cudaMemcpyHostToDevice, h->stream));
    rc = allreduce ? nngp_loglik_device_allreduce(h, kernel_id, h->d_params, K, h->d_out, h->stream)
                   : nngp_loglik_device(h, kernel_id, h->d_params, K, h->d_out, h->stream);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(ho, h->d_out, sizeof(double) * NNGP_NSTAT * K, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    memcpy(out, ho, sizeof(double) * NNGP_NSTAT * K);
    return NNGP_OK;
}

int nngp_loglik(nngp_handle *h, int kernel_id, const double *params, int K, double *out)
{
    return loglik_host(h, kernel_id, params, K, out, false);
}

int nngp_loglik_allreduce(nngp_handle *h, int kernel_id, const double *params, int K, double *out)
{
    return loglik_host(h, kernel_id, params, K, out, true);
}

// shared by nngp_factors / nngp_cov_blocks: run the emitting variant over [i0, i1) in slabs
static int run_emit(nngp_handle *h, int kernel_id, const double *params, int64_t i0, int64_t i1, double *B,
                    double *F, double *CN, double *cc, double *cs)
{
    int rc = check_eval(h, kernel_id, params, 1);
    if (rc) return rc;
    if (i0 < 0 || i1 < i0 || i1 > h->n) return fail(h, NNGP_EINVAL, "need 0 <= i0 <= i1 <= n");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int m = h->m;
    const int64_t slab = CN ? 65536 : 1 << 20;
    const int64_t cap = (i1 - i0) < slab ? (i1 - i0) : slab;
    if (cap == 0) return NNGP_OK;
    double *dB = nullptr, *dF = nullptr, *dCN = nullptr, *dcc = nullptr, *dcs = nullptr;
    auto cleanup = [&]() { free_dev(dB); free_dev(dF); free_dev(dCN); free_dev(dcc); free_dev(dcs); };
#define EMIT_TRY(call)                                                         \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) { cleanup(); return cuda_fail(h, e_, #call); }  \
    } while (0)
    if (B) EMIT_TRY(cudaMalloc(&dB, sizeof(double) * cap * m));
    if (F) EMIT_TRY(cudaMalloc(&dF, sizeof(double) * cap));
    if (CN) EMIT_TRY(cudaMalloc(&dCN, sizeof(double) * cap * m * m));
    if (cc) EMIT_TRY(cudaMalloc(&dcc, sizeof(double) * cap * m));
    if (cs) EMIT_TRY(cudaMalloc(&dcs, sizeof(double) * cap));
    if ((rc = ensure_scratch(h, 1, grid_for(h, kernel_id, cap)))) { cleanup(); return rc; }
    memcpy(h->h_stage, params, sizeof(double) * NNGP_NPARAM);
    EMIT_TRY(cudaMemcpyAsync(h->d_params, h->h_stage, sizeof(double) * NNGP_NPARAM, cudaMemcpyHostToDevice, h->stream));
    for (int64_t s0 = i0; s0 < i1; s0 += cap) {
        const int64_t s1 = s0 + cap < i1 ? s0 + cap : i1;
        const int64_t cnt = s1 - s0;
        EvalArgs a{};
        a.pts = h->pts; a.eps2 = h->eps2; a.nbr = h->nbr;
        a.lo = s0; a.hi = s1; a.m = m;
        a.params = h->d_params; a.partials = h->d_partials; a.counters = h->d_counters; a.out = h->d_out;
        a.emit = 1; a.exp2tab = h->d_exp2tab; a.K = 1; a.B = dB; a.F = dF; a.CN = dCN; a.cc = dcc; a.cs = dcs;
        if (dCN) EMIT_TRY(cudaMemsetAsync(dCN, 0, sizeof(double) * cnt * m * m, h->stream));
        if (dcc) EMIT_TRY(cudaMemsetAsync(dcc, 0, sizeof(double) * cnt * m, h->stream));
        EMIT_TRY(family_launch(h->dtype, kernel_id, m, h->D, a, 1, grid_for(h, kernel_id, cnt), h->stream));
        ++h->launches;
        const int64_t o = s0 - i0;
        if (B) EMIT_TRY(cudaMemcpyAsync(B + o * m, dB, sizeof(double) * cnt * m, cudaMemcpyDeviceToHost, h->stream));
        if (F) EMIT_TRY(cudaMemcpyAsync(F + o, dF, sizeof(double) * cnt, cudaMemcpyDeviceToHost, h->stream));
        if (CN) EMIT_TRY(cudaMemcpyAsync(CN + o * m * m, dCN, sizeof(double) * cnt * m * m, cudaMemcpyDeviceToHost, h->stream));
        if (cc) EMIT_TRY(cudaMemcpyAsync(cc + o * m, dcc, sizeof(double) * cnt * m, cudaMemcpyDeviceToHost, h->stream));
        if (cs) EMIT_TRY(cudaMemcpyAsync(cs + o, dcs, sizeof(double) * cnt, cudaMemcpyDeviceToHost, h->stream));
        EMIT_TRY(cudaStreamSynchronize(h->stream));
    }
#undef EMIT_TRY
    cleanup();
    return NNGP_OK;
}

int nngp_factors(nngp_handle *h, int kernel_id, const double *params, int64_t i0, int64_t i1, double *B, double *F)
{
    return run_emit(h, kernel_id, params, i0, i1, B, F, nullptr, nullptr, nullptr);
}

int nngp_cov_blocks(nngp_handle *h, int kernel_id, const double *params, int64_t i0, int64_t i1, double *CN,
                    double *cc, double *cs)
{
    return run_emit(h, kernel_id, params, i0, i1, nullptr, nullptr, CN, cc, cs);
}

int64_t nngp_launch_count(const nngp_handle *h) { return h ? h->launches : 0; }

int nngp_measure_fma_peak(nngp_handle *h, int dtype, int iters, double *instr_per_s)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!instr_per_s || iters < 1) return fail(h, NNGP_EINVAL, "bad arguments");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, launch_fma_peak(h, dtype, iters, instr_per_s));
    return NNGP_OK;
}

}  // extern "C"
