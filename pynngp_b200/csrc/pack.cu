// pack.cu -- upload side of nngp_set_data: the caller's row-major coordinates / y / eps2 arrive as raw
// arrays and are packed on the device into the 32-byte {x, y, z, yval} records of the hot path (one L2
// sector per neighbour gather; z carries eps2 when D < 3).  The same pass reduces the bounding box and a
// finiteness flag for the grid search of stage 1.  Replaces a host loop over n records (0.18 s at n = 1e7).
#include <math.h>

#include "nngp_common.cuh"

namespace nngp_pack {

constexpr int kBlock = 256;

// partials: gridDim.x x 7 doubles = {min x, y, z, max x, y, z, non-finite count}
__global__ void __launch_bounds__(kBlock) pack_records_kernel(const double *__restrict__ coords, const double *__restrict__ y,
                                                               const double *__restrict__ eps2, int64_t n, int D,
                                                               double4 *__restrict__ pts, double *__restrict__ partials)
{
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    double bad = 0.0;
    for (int64_t i = blockIdx.x * int64_t(kBlock) + threadIdx.x; i < n; i += int64_t(gridDim.x) * kBlock) {
        double c[3] = {0.0, 0.0, 0.0};
        for (int d = 0; d < D; ++d) {
            c[d] = coords[i * D + d];
            if (!isfinite(c[d])) bad += 1.0;
            lo[d] = fmin(lo[d], c[d]);  // fmin/fmax drop NaNs; the flag above records them
            hi[d] = fmax(hi[d], c[d]);
        }
        if (D < 3) c[2] = eps2 ? eps2[i] : 0.0;
        pts[i] = make_double4(c[0], c[1], c[2], y[i]);
    }
    __shared__ double red[kBlock / 32][7];
    double v[7] = {lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], bad};
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[k] = fmin(v[k], __shfl_xor_sync(0xffffffffu, v[k], off));
            v[3 + k] = fmax(v[3 + k], __shfl_xor_sync(0xffffffffu, v[3 + k], off));
        }
        v[6] += __shfl_xor_sync(0xffffffffu, v[6], off);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
        for (int k = 0; k < 7; ++k) red[warp][k] = v[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kBlock / 32; ++w) {
            for (int k = 0; k < 3; ++k) {
                red[0][k] = fmin(red[0][k], red[w][k]);
                red[0][3 + k] = fmax(red[0][3 + k], red[w][3 + k]);
            }
            red[0][6] += red[w][6];
        }
        for (int k = 0; k < 7; ++k) partials[size_t(blockIdx.x) * 7 + k] = red[0][k];
    }
}

// nngp_set_y: the new response arrives as one contiguous copy and is written into the records' yval lane here
// (a strided 8-byte-per-row cudaMemcpy2D of the same data is n separate DMA rows: milliseconds at n = 1e6)
__global__ void __launch_bounds__(kBlock) scatter_lane_kernel(const double *__restrict__ v, int64_t n, int lane,
                                                               double4 *__restrict__ pts)
{
    for (int64_t i = blockIdx.x * int64_t(kBlock) + threadIdx.x; i < n; i += int64_t(gridDim.x) * kBlock)
        reinterpret_cast<double *>(pts + i)[lane] = v[i];
}

// one thread per row of an injected neighbour table: entries in [-1, i), padding at the tail only
__global__ void __launch_bounds__(kBlock) validate_table_kernel(const int32_t *__restrict__ rows, int m, int64_t i0, int64_t i1,
                                                                 int32_t *viol)
{
    int bad = 0;
    for (int64_t i = i0 + blockIdx.x * int64_t(kBlock) + threadIdx.x; i < i1; i += int64_t(gridDim.x) * kBlock) {
        const int32_t *row = rows + (i - i0) * m;
        bool pad = false, ok = true;
        for (int k = 0; k < m; ++k) {
            const int32_t e = row[k];
            if (e == -1) pad = true;
            else if (e < -1 || e >= i || pad) ok = false;
        }
        bad += !ok;
    }
    if (bad) atomicAdd(viol, bad);
}

}  // namespace nngp_pack

cudaError_t launch_validate_table(nngp_handle *h, const int32_t *rows, int m, int64_t i0, int64_t i1, int32_t *d_viol,
                                  cudaStream_t stream)
{
    using namespace nngp_pack;
    cudaError_t e = cudaMemsetAsync(d_viol, 0, sizeof(int32_t), stream);
    if (e != cudaSuccess) return e;
    int grid = int(std::min<int64_t>((i1 - i0 + kBlock - 1) / kBlock, int64_t(h->num_sms) * 8));
    if (grid < 1) grid = 1;
    validate_table_kernel<<<grid, kBlock, 0, stream>>>(rows, m, i0, i1, d_viol);
    ++h->launches;
    return cudaGetLastError();
}

cudaError_t scratch_get(nngp_handle *h, int slot, size_t bytes, void **p)
{
    if (h->knn_scratch_bytes[slot] < bytes) {
        free_dev_on(h, h->knn_scratch[slot]);
        h->knn_scratch[slot] = nullptr;
        h->knn_scratch_bytes[slot] = 0;
        const size_t want = bytes + bytes / 8;  // a little headroom: the next build is often slightly larger
        cudaError_t e = dev_malloc_on(h, &h->knn_scratch[slot], want);
        if (e != cudaSuccess) return e;
        h->knn_scratch_bytes[slot] = want;
    }
    *p = h->knn_scratch[slot];
    return cudaSuccess;
}

cudaError_t launch_scatter_lane(nngp_handle *h, const double *d_v, int lane, cudaStream_t stream)
{
    using namespace nngp_pack;
    int grid = int(std::min<int64_t>((h->n + kBlock - 1) / kBlock, int64_t(h->num_sms) * 8));
    if (grid < 1) grid = 1;
    scatter_lane_kernel<<<grid, kBlock, 0, stream>>>(d_v, h->n, lane, h->pts);
    ++h->launches;
    return cudaGetLastError();
}

// d_coords (n x D), d_y (n), d_eps2 (n or null) are device copies of the caller's arrays; fills h->pts and
// the handle's bounding box / finiteness flag.
cudaError_t launch_pack_records(nngp_handle *h, const double *d_coords, const double *d_y, const double *d_eps2,
                                cudaStream_t stream)
{
    using namespace nngp_pack;
    const int64_t n = h->n;
    int grid = int(std::min<int64_t>((n + kBlock - 1) / kBlock, int64_t(h->num_sms) * 4));
    if (grid < 1) grid = 1;
    double *d_part = nullptr;
    cudaError_t e = dev_malloc_on(h, &d_part, sizeof(double) * 7 * size_t(grid));
    if (e != cudaSuccess) return e;
    pack_records_kernel<<<grid, kBlock, 0, stream>>>(d_coords, d_y, d_eps2, n, h->D, h->pts, d_part);
    ++h->launches;
    std::vector<double> part(size_t(grid) * 7);
    if ((e = cudaGetLastError()) == cudaSuccess)
        e = cudaMemcpyAsync(part.data(), d_part, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    free_dev_on(h, d_part);
    if (e != cudaSuccess) return e;
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, bad = 0.0;
    for (int b = 0; b < grid; ++b) {
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], part[size_t(b) * 7 + k]);
            hi[k] = std::max(hi[k], part[size_t(b) * 7 + 3 + k]);
        }
        bad += part[size_t(b) * 7 + 6];
    }
    for (int d = 0; d < 3; ++d) {
        h->bb_lo[d] = d < h->D ? lo[d] : 0.0;
        h->bb_hi[d] = d < h->D ? hi[d] : 0.0;
    }
    h->bb_finite = bad == 0.0;
    return cudaSuccess;
}
