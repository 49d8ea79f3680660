// nngp_api.cu -- the C ABI of libnngp_b200.so (include/nngp_b200.h): handles, uploads, dispatch.
// No CPU fallback exists anywhere in this library: every compute entry point launches a kernel.
//
// Two kinds of handle sit behind the one opaque type:
//   - a device handle (nngp_create): one CUDA device, one shard of the ordering;
//   - a group handle (nngp_create_multi): one device handle per listed device inside ONE process, peers
//     mapped with cudaDeviceEnablePeerAccess, each driven by its own host thread (nngp_group.cuh) so the
//     devices' kernels are launched in parallel; every entry point fans out over the sub-handles.
// Host-pointer evaluations (nngp_loglik, nngp_loglik_terms, ...) carry up to NNGP_PV_MAX parameter vectors
// in the kernel arguments and receive the statistics as stamped lines the kernel's last block stores into
// mapped pinned host memory -- no H2D copy, no D2H copy, no stream synchronisation on that path.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <cmath>
#include <mutex>
#include <vector>

#include "nngp_common.cuh"
#include "nngp_group.cuh"

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

// every caller has its handle in `h`
#define free_dev(p) free_dev_on(h, p)

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

// Parameter-independent part of the covariance build's exp table, one per device for the life of the process:
// 2^(j/2048), j < 2048, followed by the table the unrolled kernels copy into shared memory as it is -- 2^(j/256),
// j < 256, in 16 copies (copy c of entry j at [2048 + j * 16 + c]: one copy per lane of a half-warp, see
// loglik_fused.cuh).  (Built per handle it was an allocation, a copy and a stream synchronisation in every nngp_create.)
std::mutex g_tab_mu;
double *g_exp2tab[64] = {};

cudaError_t shared_exp2tab(int device, double **out)
{
    std::lock_guard<std::mutex> lk(g_tab_mu);
    if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
    if (!g_exp2tab[device]) {
        std::vector<double> tab(2048 + 4096);
        for (int j = 0; j < 2048; ++j) tab[j] = exp2(double(j) / 2048.0);
        for (int j = 0; j < 256; ++j)
            for (int c = 0; c < 16; ++c) tab[2048 + j * 16 + c] = tab[j * 8];
        double *d = nullptr;
        cudaError_t e = cudaMalloc(&d, sizeof(double) * tab.size());
        if (e == cudaSuccess) e = cudaMemcpy(d, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { if (d) cudaFree(d); return e; }
        g_exp2tab[device] = d;
    }
    *out = g_exp2tab[device];
    return cudaSuccess;
}

// Pinned host blocks of destroyed handles (parameter staging + the mapped result lines), kept for the next handle:
// cudaMallocHost / cudaHostAlloc cost about a millisecond each, which is most of a new handle's first evaluation.
struct HostBlock { double *stage; uint4 *out; int cap; };
std::mutex g_host_mu;
std::vector<HostBlock> g_host_free;

void host_block_put(nngp_handle *h)
{
    if (!h->h_stage && !h->h_out) return;
    std::lock_guard<std::mutex> lk(g_host_mu);
    if (h->h_stage && h->h_out && g_host_free.size() < 64) g_host_free.push_back(HostBlock{h->h_stage, h->h_out, h->K_cap});
    else {
        if (h->h_stage) cudaFreeHost(h->h_stage);
        if (h->h_out) cudaFreeHost(h->h_out);
    }
    h->h_stage = nullptr; h->h_out = nullptr;
}

cudaError_t host_block_get(nngp_handle *h, int cap)
{
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        int best = -1;
        for (int k = 0; k < int(g_host_free.size()); ++k)
            if (g_host_free[k].cap >= cap && (best < 0 || g_host_free[k].cap < g_host_free[best].cap)) best = k;
        if (best >= 0) {
            h->h_stage = g_host_free[best].stage; h->h_out = g_host_free[best].out;
            g_host_free.erase(g_host_free.begin() + best);
        }
    }
    if (!h->h_out) {
        cudaError_t e = cudaMallocHost(&h->h_stage, sizeof(double) * NNGP_NPARAM * cap);
        if (e == cudaSuccess) e = cudaHostAlloc(&h->h_out, sizeof(uint4) * NNGP_NSTAT * cap, cudaHostAllocMapped | cudaHostAllocPortable);
        if (e != cudaSuccess) return e;
    }
    memset(h->h_out, 0, sizeof(uint4) * NNGP_NSTAT * cap);  // stamp 0 = never written (a handle's stamps start at 1)
    return cudaSuccess;
}

// Evaluation scratch.  The ticket counters are zeroed on `st`, the stream the kernel is launched on: a
// memset on another stream (the legacy NULL stream included) is not ordered against a non-blocking stream.
int ensure_scratch(nngp_handle *h, int K, int grid, cudaStream_t st)
{
    if (K > h->K_cap) {
        CUDA_TRY(h, quiesce(h));  // nothing may still read the buffers being replaced
        free_dev(h->d_params); free_dev(h->d_out); free_dev(h->d_counters);
        host_block_put(h);
        int cap = K < 16 ? 16 : K;
        CUDA_TRY(h, dev_malloc_on(h, &h->d_params, sizeof(double) * NNGP_NPARAM * cap));
        CUDA_TRY(h, dev_malloc_on(h, &h->d_out, sizeof(double) * NNGP_NSTAT * cap));
        CUDA_TRY(h, dev_malloc_on(h, &h->d_counters, sizeof(unsigned int) * cap));
        if (st != h->stream) CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // the blocks were allocated in h->stream's order
        CUDA_TRY(h, cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned int) * cap, st));
        CUDA_TRY(h, host_block_get(h, cap));
        free_dev(h->d_partials);
        h->grid_cap = 0;
        h->K_cap = cap;
    }
    if (grid > h->grid_cap || !h->d_partials) {
        CUDA_TRY(h, quiesce(h));
        free_dev(h->d_partials);
        int gc = grid < 1024 ? 1024 : grid;
        CUDA_TRY(h, dev_malloc_on(h, &h->d_partials, sizeof(double) * 3 * size_t(gc) * h->K_cap));
        if (st != h->stream) CUDA_TRY(h, cudaStreamSynchronize(h->stream));
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
    if (h->hi > h->lo && !h->holds_rows(h->lo, h->hi))
        return fail(h, NNGP_ESTATE, "the neighbour table does not hold the rows of the shard (it was built for another shard)");
    return NNGP_OK;
}

// Launches one evaluation of the handle's shard.  d_params: K x 4 in device memory, or NULL with the vectors
// in pv (K <= NNGP_PV_MAX).  The K x 3 statistics go to d_out (device) and / or hout (stamped lines in mapped
// host memory, stamp `seq`); with `px` non-NULL they are the sums over all ranks of the exchange.
int launch_eval(nngp_handle *h, int kernel_id, const double *d_params, const double *pv, int K, double *d_out, uint4 *hout,
                unsigned int seq, const PeerExchange *px, cudaStream_t st)
{
    const int64_t nloc = h->hi - h->lo;
    if (nloc == 0) {  // empty shard: the statistics are exactly zero (it still takes part in an exchange)
        if (px) {
            CUDA_TRY(h, launch_peer_zero(h, *px, K, d_out, hout, seq, st));
            ++h->launches;
        } else {
            if (d_out) CUDA_TRY(h, cudaMemsetAsync(d_out, 0, sizeof(double) * NNGP_NSTAT * K, st));
            if (hout) {
                PeerExchange solo{};  // world 0: publishes zeros without an exchange
                CUDA_TRY(h, launch_peer_zero(h, solo, K, nullptr, hout, seq, st));
                ++h->launches;
            }
        }
        return NNGP_OK;
    }
    const int grid = grid_for(h, kernel_id, nloc);
    int rc = ensure_scratch(h, K, grid, st);
    if (rc) return rc;
    EvalArgs a{};
    a.pts = h->pts; a.eps2 = h->eps2; a.nbr = h->nbr_base();
    a.lo = h->lo; a.hi = h->hi; a.m = h->m;
    a.params = d_params;
    if (!d_params) memcpy(a.pv, pv, sizeof(double) * NNGP_NPARAM * K);
    a.partials = h->d_partials; a.counters = h->d_counters; a.out = d_out; a.hout = hout; a.seq = seq;
    a.emit = 0; a.exp2tab = h->d_exp2tab; a.K = K;
    a.gather_bypass_l1 = h->n * int64_t(sizeof(double4)) > (int64_t(96) << 20);  // records beyond ~ the L2's size
    if (px) a.px = *px;
    if (h->timing) CUDA_TRY(h, cudaEventRecord(h->ev0, st));
    CUDA_TRY(h, family_launch(h->dtype, kernel_id, h->m, h->D, a, K, grid, st));
    if (h->timing) CUDA_TRY(h, cudaEventRecord(h->ev1, st));
    ++h->launches;
    return NNGP_OK;
}

// The host side of the stamped-line result: polls the K x 3 lines for stamp `seq` and decodes them.  The
// stream is queried now and then so that a failed launch surfaces as an error instead of an endless spin.
int wait_lines(nngp_handle *h, cudaStream_t st, int K, unsigned int seq, double *out)
{
    const volatile uint64_t *lines = reinterpret_cast<const volatile uint64_t *>(h->h_out);
    bool drained = false;
    for (int i = 0; i < K * NNGP_NSTAT; ++i) {
        uint64_t a, b;
        for (uint32_t spin = 1;; ++spin) {
            a = lines[2 * i]; b = lines[2 * i + 1];
            if (uint32_t(a >> 32) == seq && uint32_t(b >> 32) == seq) break;
            nngp_cpu_relax();
            if ((spin & 0xfffu) == 0) {
                if (drained) return fail(h, NNGP_ECUDA, "the evaluation finished without publishing its result");
                const cudaError_t q = cudaStreamQuery(st);
                if (q == cudaSuccess) drained = true;  // one more look at the line, then give up
                else if (q != cudaErrorNotReady) return cuda_fail(h, q, "evaluation kernel");
            }
        }
        const uint64_t bits = (b << 32) | (a & 0xffffffffull);
        memcpy(out + i, &bits, sizeof(double));
    }
    return NNGP_OK;
}

int next_seq(nngp_handle *h)
{
    if (++h->seq == 0) h->seq = 1;  // 0 is the 'never written' stamp
    return int(h->seq);
}

// host-pointer evaluation of one device handle
int loglik_host(nngp_handle *h, int kernel_id, const double *params, int K, double *out, bool allreduce)
{
    int rc = check_eval(h, kernel_id, params, K);
    if (rc) return rc;
    if (!out) return fail(h, NNGP_EINVAL, "out must not be NULL");
    if (allreduce) {
        if (h->px.world < 2) return fail(h, NNGP_ESTATE, "peer exchange is not connected (nngp_peer_export / nngp_peer_connect)");
        if (K > h->px.K_cap) return fail(h, NNGP_EINVAL, "K exceeds the exchange buffer's K_cap");
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    if ((rc = ensure_scratch(h, K, 1, h->stream))) return rc;
    const double *d_params = nullptr;
    if (K > NNGP_PV_MAX) {
        memcpy(h->h_stage, params, sizeof(double) * NNGP_NPARAM * K);
        CUDA_TRY(h, cudaMemcpyAsync(h->d_params, h->h_stage, sizeof(double) * NNGP_NPARAM * K, cudaMemcpyHostToDevice, h->stream));
        d_params = h->d_params;
    }
    const unsigned int seq = (unsigned int)next_seq(h);
    PeerExchange px = h->px;
    px.gen = h->px.gen + 1;  // every rank issues the same sequence of exchanges
    if ((rc = launch_eval(h, kernel_id, d_params, params, K, nullptr, h->h_out, seq, allreduce ? &px : nullptr, h->stream))) return rc;
    if (allreduce) h->px.gen = px.gen;  // only a launch that happened advances the generation
    return wait_lines(h, h->stream, K, seq, out);
}

// ---- group handles ------------------------------------------------------------------------------------
bool is_group(const nngp_handle *h) { return h && h->group; }

// contiguous block r of [lo, hi) split over `world`
void split_rows(int64_t lo, int64_t hi, int r, int world, int64_t *a, int64_t *b)
{
    const int64_t len = hi - lo;
    *a = lo + (len * r) / world;
    *b = lo + (len * (r + 1)) / world;
}

// runs f(sub, r) for every sub-handle on its own thread; the first failure becomes the group's error
template <class F>
int group_each(nngp_handle *h, F &&f)
{
    nngp_group *g = h->group;
    const int rc = g->run([&](int r) { return f(g->subs[r], r); });
    if (rc) {
        for (nngp_handle *s : g->subs)
            if (!s->err.empty()) { h->err = s->err; s->err.clear(); break; }
    }
    return rc;
}

int group_unsupported(nngp_handle *h, const char *what)
{
    return fail(h, NNGP_ESTATE, std::string(what) + " is not available on a multi-device handle (nngp_create_multi)");
}

}  // namespace

extern "C" {

const char *nngp_version(void) { return "nngp_b200 0.2 (sm_100a)"; }

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
    // (single attributes: cudaGetDeviceProperties takes milliseconds)
    int cc_major = 0, cc_minor = 0, num_sms = 0;
    if ((e = cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, device)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess)
        return cuda_fail(nullptr, e, "cudaDeviceGetAttribute");
    if (cc_major != 10)
        return fail(nullptr, NNGP_ENODEVICE, "libnngp_b200 is built for sm_100a only; device is sm_" +
                                                 std::to_string(cc_major) + std::to_string(cc_minor));
    if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    nngp_handle *h = new nngp_handle();
    h->device = device;
    h->dtype = dtype;
    h->num_sms = num_sms;
    auto bail = [&](cudaError_t err, const char *what) {
        const int rc = cuda_fail(nullptr, err, what);
        nngp_destroy(h);
        return rc;
    };
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    {   // freed blocks stay in the device's pool (up to 8 GB) instead of going back to the driver at every synchronisation
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            uint64_t keep = 8ull << 30, cur = 0;
            if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &cur) == cudaSuccess && cur < keep)
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    if ((e = dev_malloc_on(h, &h->d_tile_counter, sizeof(unsigned int))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = dev_malloc_on(h, &h->d_viol, sizeof(int32_t))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = shared_exp2tab(device, &h->d_exp2tab)) != cudaSuccess) return bail(e, "cudaMalloc (exp table)");
    *out = h;
    return NNGP_OK;
}

int nngp_create_multi(nngp_handle **out, const int *devices, int ndev, int dtype)
{
    if (!out) return fail(nullptr, NNGP_EINVAL, "null handle pointer");
    *out = nullptr;
    if (!devices || ndev < 1 || ndev > NNGP_MAX_PEERS)
        return fail(nullptr, NNGP_EINVAL, "need 1 <= ndev <= 8 device indices");
    for (int a = 0; a < ndev; ++a)
        for (int b = 0; b < a; ++b)
            if (devices[a] == devices[b]) return fail(nullptr, NNGP_EINVAL, "a device is listed twice");
    nngp_handle *h = new nngp_handle();
    h->dtype = dtype;
    nngp_group *g = h->group = new nngp_group();
    for (int r = 0; r < ndev; ++r) {
        nngp_handle *s = nullptr;
        const int rc = nngp_create(&s, devices[r], dtype);
        if (rc) { nngp_destroy(h); return rc; }  // the message is already in g_create_error
        g->subs.push_back(s);
    }
    h->device = devices[0];
    h->num_sms = g->subs[0]->num_sms;
    // every device maps every other device's memory (NVLink / NVSwitch peer access): the exchange buffers are
    // plain device pointers inside one process
    for (int a = 0; a < ndev && ndev > 1; ++a) {
        cudaSetDevice(devices[a]);
        for (int b = 0; b < ndev; ++b) {
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
            cudaError_t e = can ? cudaDeviceEnablePeerAccess(devices[b], 0) : cudaErrorPeerAccessUnsupported;
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e != cudaSuccess) {
                const int rc = cuda_fail(nullptr, e, "cudaDeviceEnablePeerAccess (a multi-device handle needs P2P access between all its devices)");
                nngp_destroy(h);
                return rc;
            }
        }
    }
    if (ndev > 1) {
        const int K_cap = 128;
        std::vector<uint4 *> bufs(ndev, nullptr);
        for (int r = 0; r < ndev; ++r) {
            nngp_handle *s = g->subs[r];
            cudaSetDevice(s->device);
            const size_t bytes = sizeof(uint4) * 2 * NNGP_MAX_PEERS * size_t(K_cap) * 3;
            cudaError_t e = cudaMalloc(&s->xbuf, bytes);
            if (e == cudaSuccess) e = cudaMemset(s->xbuf, 0, bytes);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { const int rc = cuda_fail(nullptr, e, "cudaMalloc (exchange buffer)"); nngp_destroy(h); return rc; }
            bufs[r] = static_cast<uint4 *>(s->xbuf);
        }
        for (int r = 0; r < ndev; ++r) {
            nngp_handle *s = g->subs[r];
            s->px.rank = r; s->px.world = ndev; s->px.K_cap = K_cap; s->px.gen = 0;
            for (int q = 0; q < ndev; ++q) s->px.lines[q] = bufs[q];
        }
    }
    g->start();
    *out = h;
    return NNGP_OK;
}

int nngp_device_count(const nngp_handle *h) { return !h ? 0 : h->group ? int(h->group->subs.size()) : 1; }

int nngp_visible_devices(void)
{
    int ndev = 0, ok = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); return 0; }
    for (int d = 0; d < ndev; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
        else break;  // device indices handed to nngp_create_multi are 0..count-1
    }
    return ok;
}

void nngp_destroy(nngp_handle *h)
{
    if (!h) return;
    if (h->group) {
        h->group->stop();
        for (nngp_handle *s : h->group->subs) nngp_destroy(s);
        delete h->group;
        delete h;
        return;
    }
    cudaSetDevice(h->device);
    if (h->stream) quiesce(h);
    free_dev(h->pts); free_dev(h->eps2); free_dev(h->nbr); free_dev(h->d_ystage);
    free_dev(h->d_params); free_dev(h->d_out); free_dev(h->d_partials);
    free_dev(h->d_counters); free_dev(h->d_tile_counter); free_dev(h->d_viol);
    h->d_exp2tab = nullptr;  // the device's shared table (shared_exp2tab) outlives the handle
    for (void *&p : h->knn_scratch) free_dev(p);
    for (int r = 0; r < NNGP_MAX_PEERS; ++r)
        if (h->peer_base[r] && h->peer_base[r] != h->xbuf) cudaIpcCloseMemHandle(h->peer_base[r]);
    if (h->xbuf) cudaFree(h->xbuf);
    h->xbuf = nullptr;
    host_block_put(h);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int nngp_set_data(nngp_handle *h, const double *coords, int64_t n, int D, const double *y, const double *eps2)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!coords || !y) return fail(h, NNGP_EINVAL, "coords and y must not be NULL");
    if (n < 1 || n > 2147483647LL) return fail(h, NNGP_EINVAL, "n must be in [1, 2^31)");
    if (D < 1 || D > NNGP_MAX_D) return fail(h, NNGP_EINVAL, "D must be 1, 2 or 3");
    if (is_group(h)) {
        // replicated on every device (each over its own PCIe link, in parallel); the shard [0, n) is split evenly
        int rc = group_each(h, [&](nngp_handle *s, int) { return nngp_set_data(s, coords, n, D, y, eps2); });
        if (rc) return rc;
        h->n = n; h->D = D; h->m = 0; h->has_nbr = false;
        return nngp_set_shard(h, 0, n);
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, quiesce(h));
    free_dev(h->pts); free_dev(h->eps2); free_dev(h->nbr); free_dev(h->d_ystage);  // the landing buffer is sized by n
    h->has_nbr = false; h->m = 0; h->nbr_row0 = 0; h->nbr_rows = 0;
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
    SET_TRY(dev_malloc_on(h, &h->pts, sizeof(double4) * (size_t)n));
    SET_TRY(dev_malloc_on(h, &d_coords, sizeof(double) * (size_t)n * D));
    SET_TRY(dev_malloc_on(h, &d_y, sizeof(double) * (size_t)n));
    SET_TRY(cudaMemcpyAsync(d_coords, coords, sizeof(double) * (size_t)n * D, cudaMemcpyHostToDevice, h->stream));
    SET_TRY(cudaMemcpyAsync(d_y, y, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    if (eps2) {
        SET_TRY(dev_malloc_on(h, &d_e2, sizeof(double) * (size_t)n));
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
    if (is_group(h)) {
        if (!h->n) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
        return group_each(h, [&](nngp_handle *s, int) { return nngp_set_y(s, y); });
    }
    if (!h->pts) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
    if (!y) return fail(h, NNGP_EINVAL, "y must not be NULL");
    CUDA_TRY(h, cudaSetDevice(h->device));
    // one contiguous copy, then a kernel writes the yval lane of the records (pack.cu)
    if (!h->d_ystage) CUDA_TRY(h, dev_malloc_on(h, &h->d_ystage, sizeof(double) * (size_t)h->n));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_ystage, y, sizeof(double) * (size_t)h->n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, launch_scatter_lane(h, h->d_ystage, 3, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NNGP_OK;
}

int nngp_set_eps2(nngp_handle *h, const double *eps2)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (is_group(h)) {
        if (!h->n) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
        return group_each(h, [&](nngp_handle *s, int) { return nngp_set_eps2(s, eps2); });
    }
    if (!h->pts) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
    if (!eps2) return fail(h, NNGP_EINVAL, "eps2 must not be NULL");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->D > 2) {  // D = 3: its own array (the records' z slot holds a coordinate)
        if (!h->eps2) CUDA_TRY(h, dev_malloc_on(h, &h->eps2, sizeof(double) * (size_t)h->n));
        CUDA_TRY(h, cudaMemcpyAsync(h->eps2, eps2, sizeof(double) * (size_t)h->n, cudaMemcpyHostToDevice, h->stream));
    } else {         // D < 3: the records' z slot
        if (!h->d_ystage) CUDA_TRY(h, dev_malloc_on(h, &h->d_ystage, sizeof(double) * (size_t)h->n));
        CUDA_TRY(h, cudaMemcpyAsync(h->d_ystage, eps2, sizeof(double) * (size_t)h->n, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, launch_scatter_lane(h, h->d_ystage, 2, h->stream));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NNGP_OK;
}

int nngp_set_shard(nngp_handle *h, int64_t lo, int64_t hi)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (is_group(h)) {
        if (!h->n) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
        if (lo < 0 || hi < lo || hi > h->n) return fail(h, NNGP_EINVAL, "shard must satisfy 0 <= lo <= hi <= n");
        nngp_group *g = h->group;
        const int world = int(g->subs.size());
        for (int r = 0; r < world; ++r) {
            int64_t a, b;
            split_rows(lo, hi, r, world, &a, &b);
            const int rc = nngp_set_shard(g->subs[r], a, b);
            if (rc) { h->err = g->subs[r]->err; return rc; }
        }
        h->lo = lo; h->hi = hi;
        return NNGP_OK;
    }
    if (!h->pts) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
    if (lo < 0 || hi < lo || hi > h->n) return fail(h, NNGP_EINVAL, "shard must satisfy 0 <= lo <= hi <= n");
    h->lo = lo; h->hi = hi;
    return NNGP_OK;
}

// (re)allocates the table for rows [row0, row0 + rows) of the n x m table
static int alloc_nbr(nngp_handle *h, int m, int64_t row0, int64_t rows)
{
    if (m < 1 || m > NNGP_MAX_M) return fail(h, NNGP_EINVAL, "m must be in [1, 32]");
    if (h->nbr && (h->m != m || h->nbr_rows != rows)) {
        CUDA_TRY(h, quiesce(h));
        free_dev(h->nbr);
    }
    if (!h->nbr) CUDA_TRY(h, dev_malloc_on(h, &h->nbr, sizeof(int32_t) * (size_t)(rows > 0 ? rows : 1) * m));
    h->m = m; h->nbr_row0 = row0; h->nbr_rows = rows;
    h->has_nbr = false;
    return NNGP_OK;
}

int nngp_build_neighbors(nngp_handle *h, int m, int tile_offset, int tile_stride)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (is_group(h)) return group_unsupported(h, "nngp_build_neighbors (use nngp_build_neighbors_grid with NNGP_KNN_BRUTE)");
    if (!h->pts) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
    if (tile_stride < 1 || tile_offset < 0 || tile_offset >= tile_stride)
        return fail(h, NNGP_EINVAL, "need 0 <= tile_offset < tile_stride");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = alloc_nbr(h, m, 0, h->n);
    if (rc) return rc;
    CUDA_TRY(h, launch_knn_ordered(h, m, tile_offset, tile_stride, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->has_nbr = true;
    return NNGP_OK;
}

// window = true: the table holds rows [row_lo, row_hi) only
static int build_neighbors_capped(nngp_handle *h, int m, int64_t row_lo, int64_t row_hi, int64_t cand_cap, int algo, bool window)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!h->pts) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
    if (row_lo < 0 || row_hi < row_lo || row_hi > h->n) return fail(h, NNGP_EINVAL, "need 0 <= row_lo <= row_hi <= n");
    if (cand_cap < 1) return fail(h, NNGP_EINVAL, "cand_cap must be >= 1");
    if (algo != NNGP_KNN_AUTO && algo != NNGP_KNN_GRID && algo != NNGP_KNN_BRUTE)
        return fail(h, NNGP_EINVAL, "algo must be NNGP_KNN_AUTO, NNGP_KNN_GRID or NNGP_KNN_BRUTE");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = window ? alloc_nbr(h, m, row_lo, row_hi - row_lo) : alloc_nbr(h, m, 0, h->n);
    if (rc) return rc;
    int32_t *table = h->nbr_base();
    int used = 0;
    if (row_hi > row_lo) {
        if (algo != NNGP_KNN_BRUTE) {
            CUDA_TRY(h, launch_knn_grid(h, true, m, row_lo, row_hi, cand_cap, table, window, h->stream, algo == NNGP_KNN_GRID, &used));
            if (!used && algo == NNGP_KNN_GRID) return fail(h, NNGP_EINVAL, "grid search needs finite coordinates");
        }
        if (!used) {
            // brute force over the query tiles covering [row_lo, row_hi); every other row is unset
            if (!window) CUDA_TRY(h, launch_fill_i32(h, table, row_lo * m, NNGP_ROW_UNSET, h->stream));
            CUDA_TRY(h, launch_knn_brute_rows(h, m, row_lo, row_hi, cand_cap, table, h->stream));
            if (!window) CUDA_TRY(h, launch_fill_i32(h, table + row_hi * m, (h->n - row_hi) * m, NNGP_ROW_UNSET, h->stream));
        }
    } else if (!window) {
        CUDA_TRY(h, launch_fill_i32(h, table, h->n * int64_t(m), NNGP_ROW_UNSET, h->stream));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->knn_used_grid = used;
    h->has_nbr = true;
    return NNGP_OK;
}

int nngp_build_neighbors_shard(nngp_handle *h, int m, int algo)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (is_group(h)) {
        if (!h->n) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
        const int rc = group_each(h, [&](nngp_handle *s, int) { return nngp_build_neighbors_shard(s, m, algo); });
        if (rc) return rc;
        h->m = m; h->has_nbr = true;
        return NNGP_OK;
    }
    if (!h->pts) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
    return build_neighbors_capped(h, m, h->lo, h->hi, INT64_MAX, algo, true);
}

int nngp_build_neighbors_grid(nngp_handle *h, int m, int64_t row_lo, int64_t row_hi, int algo)
{
    if (is_group(h)) {
        // every device searches the rows of its own shard and keeps only those
        if (row_lo > h->lo || row_hi < h->hi) return fail(h, NNGP_EINVAL, "a multi-device handle builds the rows of its whole shard: need row_lo <= lo and hi <= row_hi");
        return nngp_build_neighbors_shard(h, m, algo);
    }
    return build_neighbors_capped(h, m, row_lo, row_hi, INT64_MAX, algo, false);
}

int nngp_build_neighbors_capped(nngp_handle *h, int m, int64_t row_lo, int64_t row_hi, int64_t cand_cap, int algo)
{
    if (is_group(h)) return group_unsupported(h, "nngp_build_neighbors_capped");
    return build_neighbors_capped(h, m, row_lo, row_hi, cand_cap, algo, false);
}

int nngp_set_knn_tuning(nngp_handle *h, double lambda_scale, int64_t brute_rows)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!(lambda_scale > 0.0) || brute_rows < 1) return fail(h, NNGP_EINVAL, "need lambda_scale > 0 and brute_rows >= 1");
    if (is_group(h))
        for (nngp_handle *s : h->group->subs) { s->knn_lambda_scale = lambda_scale; s->knn_brute_rows = brute_rows; }
    h->knn_lambda_scale = lambda_scale;
    h->knn_brute_rows = brute_rows;
    return NNGP_OK;
}

int nngp_knn_used_grid(const nngp_handle *h)
{
    if (!h) return 0;
    if (h->group) {
        for (const nngp_handle *s : h->group->subs)
            if (s->hi > s->lo && !s->knn_used_grid) return 0;
        return 1;
    }
    return h->knn_used_grid;
}

// Uploads rows [i0, i1) of a caller's table (idx points at row i0) as this handle's table window and checks it on
// the device: every entry in [-1, i) (a neighbour precedes its row: the sets are causal and never reach past the
// table), padding only at the tail.  The fused kernel gathers pts[idx] for every entry >= 0, so a foreign or
// stale table must be refused here, not read.
static int upload_table(nngp_handle *h, const int32_t *idx, int m, int64_t i0, int64_t i1)
{
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = alloc_nbr(h, m, i0, i1 - i0);
    if (rc) return rc;
    if (i1 > i0) {
        CUDA_TRY(h, cudaMemcpyAsync(h->nbr, idx, sizeof(int32_t) * (size_t)(i1 - i0) * m, cudaMemcpyHostToDevice, h->stream));
        int32_t viol = 0;
        CUDA_TRY(h, launch_validate_table(h, h->nbr, m, i0, i1, h->d_viol, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(&viol, h->d_viol, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        if (viol)
            return fail(h, NNGP_EINVAL, std::to_string(viol) + " row(s) of the neighbour table are invalid: entries must lie in [-1, i) "
                                        "for row i (neighbours precede their row) with the -1 padding at the tail");
    }
    h->has_nbr = true;
    return NNGP_OK;
}

int nngp_set_neighbors(nngp_handle *h, const int32_t *idx, int m)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!idx) return fail(h, NNGP_EINVAL, "idx must not be NULL");
    if (is_group(h)) {
        if (!h->n) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
        if (m < 1 || m > NNGP_MAX_M) return fail(h, NNGP_EINVAL, "m must be in [1, 32]");
        const int rc = group_each(h, [&](nngp_handle *s, int) { return upload_table(s, idx + s->lo * m, m, s->lo, s->hi); });
        if (rc) return rc;
        h->m = m; h->has_nbr = true;
        return NNGP_OK;
    }
    if (!h->pts) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
    return upload_table(h, idx, m, 0, h->n);
}

int nngp_get_neighbor_rows(nngp_handle *h, int64_t i0, int64_t i1, int32_t *out)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!h->has_nbr) return fail(h, NNGP_ESTATE, "no neighbour table");
    if (!out || i0 < 0 || i1 < i0 || i1 > h->n) return fail(h, NNGP_EINVAL, "need out != NULL and 0 <= i0 <= i1 <= n");
    if (i1 == i0) return NNGP_OK;
    if (is_group(h)) {
        // rows come from the devices that hold them; rows outside the group's shard were never built
        const int m = h->m;
        for (int64_t k = 0; k < (i1 - i0) * m; ++k) out[k] = NNGP_ROW_UNSET;
        return group_each(h, [&](nngp_handle *s, int) {
            const int64_t a = std::max(i0, s->nbr_row0), b = std::min(i1, s->nbr_row0 + s->nbr_rows);
            return a < b ? nngp_get_neighbor_rows(s, a, b, out + (a - i0) * m) : NNGP_OK;
        });
    }
    if (!h->holds_rows(i0, i1)) return fail(h, NNGP_ESTATE, "the table of this handle holds the rows of its shard only");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaMemcpyAsync(out, h->nbr_base() + i0 * h->m, sizeof(int32_t) * size_t(i1 - i0) * h->m, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NNGP_OK;
}

int nngp_get_neighbors(nngp_handle *h, int32_t *out)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    return nngp_get_neighbor_rows(h, 0, h->n, out);
}

int nngp_neighbor_window(const nngp_handle *h, int64_t *row0, int64_t *rows)
{
    if (!h || !row0 || !rows) return NNGP_EINVAL;
    if (h->group) { *row0 = h->lo; *rows = h->has_nbr ? h->hi - h->lo : 0; return NNGP_OK; }
    *row0 = h->nbr_row0; *rows = h->has_nbr ? h->nbr_rows : 0;
    return NNGP_OK;
}

int nngp_knn_plain(nngp_handle *h, int k, int32_t *out)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (is_group(h)) {  // replicated coordinates: the first device answers
        const int rc = nngp_knn_plain(h->group->subs[0], k, out);
        if (rc) h->err = h->group->subs[0]->err;
        return rc;
    }
    if (!h->pts) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
    if (!out || k < 1 || k > NNGP_MAX_M) return fail(h, NNGP_EINVAL, "need out != NULL and 1 <= k <= 32");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int32_t *d_tab = nullptr;
    CUDA_TRY(h, dev_malloc_on(h, &d_tab, sizeof(int32_t) * (size_t)h->n * k));
    int used = 0;
    cudaError_t e = launch_knn_grid(h, false, k, 0, h->n, INT64_MAX, d_tab, false, h->stream, 0, &used);
    if (e == cudaSuccess && !used) e = launch_knn_plain(h, k, d_tab, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_tab, sizeof(int32_t) * (size_t)h->n * k, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    free_dev(d_tab);
    if (e != cudaSuccess) return cuda_fail(h, e, "nngp_knn_plain");
    return NNGP_OK;
}

void *nngp_neighbors_device_ptr(nngp_handle *h)
{
    return (h && !h->group && h->has_nbr && h->nbr_row0 == 0 && h->nbr_rows == h->n) ? (void *)h->nbr : nullptr;
}

void *nngp_neighbor_window_device_ptr(nngp_handle *h) { return (h && !h->group && h->has_nbr) ? (void *)h->nbr : nullptr; }

int nngp_loglik_device(nngp_handle *h, int kernel_id, const double *d_params, int K, double *d_out, void *stream)
{
    if (is_group(h)) return group_unsupported(h, "nngp_loglik_device");
    int rc = check_eval(h, kernel_id, d_params, K);
    if (rc) return rc;
    if (!d_out) return fail(h, NNGP_EINVAL, "d_out must not be NULL");
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    if (st != h->stream) h->foreign_stream_used = true;
    return launch_eval(h, kernel_id, d_params, nullptr, K, d_out, nullptr, 0, nullptr, st);
}

static size_t peer_bytes(int K_cap) { return sizeof(uint4) * 2 * NNGP_MAX_PEERS * size_t(K_cap) * 3; }

int nngp_peer_export(nngp_handle *h, int K_cap, unsigned char *handle_out)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (is_group(h)) return group_unsupported(h, "nngp_peer_export (its devices are already connected)");
    if (!handle_out || K_cap < 1) return fail(h, NNGP_EINVAL, "need handle_out != NULL and K_cap >= 1");
    // one spinning last block per parameter vector must be able to be resident on every rank at once
    if (K_cap > h->num_sms) return fail(h, NNGP_EINVAL, "K_cap must not exceed the number of SMs");
    if (h->px.world > 1) return fail(h, NNGP_ESTATE, "peer exchange is already connected");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->xbuf) cudaFree(h->xbuf);
    h->xbuf = nullptr;
    CUDA_TRY(h, cudaMalloc(&h->xbuf, peer_bytes(K_cap)));
    CUDA_TRY(h, cudaMemset(h->xbuf, 0, peer_bytes(K_cap)));
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
    if (is_group(h)) return group_unsupported(h, "nngp_peer_connect");
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
        h->px.lines[r] = reinterpret_cast<uint4 *>(base);
    }
    h->px.rank = rank;
    h->px.world = world;
    h->px.gen = 0;
    return NNGP_OK;
}

int nngp_loglik_device_allreduce(nngp_handle *h, int kernel_id, const double *d_params, int K, double *d_out, void *stream)
{
    if (is_group(h)) return group_unsupported(h, "nngp_loglik_device_allreduce");
    int rc = check_eval(h, kernel_id, d_params, K);
    if (rc) return rc;
    if (!d_out) return fail(h, NNGP_EINVAL, "d_out must not be NULL");
    if (h->px.world < 2) return fail(h, NNGP_ESTATE, "peer exchange is not connected (nngp_peer_export / nngp_peer_connect)");
    if (K > h->px.K_cap) return fail(h, NNGP_EINVAL, "K exceeds the exchange buffer's K_cap");
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    if (st != h->stream) h->foreign_stream_used = true;
    PeerExchange px = h->px;
    px.gen = h->px.gen + 1;  // every rank issues the same sequence of exchanges
    if ((rc = launch_eval(h, kernel_id, d_params, nullptr, K, d_out, nullptr, 0, &px, st))) return rc;
    h->px.gen = px.gen;      // only a launch that happened advances the generation
    return NNGP_OK;
}

// multi-device evaluation: every device's thread launches its shard's kernel (the sum over the devices is made
// inside the kernels' tails through peer memory); device 0 publishes the totals to the host
static int group_loglik(nngp_handle *h, int kernel_id, const double *params, int K, double *out)
{
    if (!h->n) return fail(h, NNGP_ESTATE, "nngp_set_data has not been called");
    if (!h->has_nbr) return fail(h, NNGP_ESTATE, "no neighbour table: call nngp_build_neighbors_grid or nngp_set_neighbors");
    if (kernel_id < 0 || kernel_id > NNGP_MATERN52) return fail(h, NNGP_EINVAL, "unknown kernel_id");
    if (!params || K < 1 || !out) return fail(h, NNGP_EINVAL, "params must hold K >= 1 parameter vectors and out must not be NULL");
    nngp_group *g = h->group;
    const int world = int(g->subs.size());
    if (world > 1 && K > g->subs[0]->px.K_cap) {  // chunks of the exchange buffer's capacity
        const int cap = g->subs[0]->px.K_cap;
        for (int k0 = 0; k0 < K; k0 += cap) {
            const int rc = group_loglik(h, kernel_id, params + size_t(k0) * NNGP_NPARAM, std::min(cap, K - k0), out + size_t(k0) * NNGP_NSTAT);
            if (rc) return rc;
        }
        return NNGP_OK;
    }
    nngp_handle *s0 = g->subs[0];
    cudaSetDevice(s0->device);
    int rc = ensure_scratch(s0, K, 1, s0->stream);
    if (rc) { h->err = s0->err; return rc; }
    const unsigned int seq = (unsigned int)next_seq(s0);
    rc = group_each(h, [&](nngp_handle *s, int r) {
        int rc2 = check_eval(s, kernel_id, params, K);
        if (rc2) return rc2;
        if (r == 0) { cudaError_t e = cudaSetDevice(s->device); if (e != cudaSuccess) return cuda_fail(s, e, "cudaSetDevice"); }
        const double *d_params = nullptr;
        if (K > NNGP_PV_MAX) {
            if ((rc2 = ensure_scratch(s, K, 1, s->stream))) return rc2;
            memcpy(s->h_stage, params, sizeof(double) * NNGP_NPARAM * K);
            cudaError_t e = cudaMemcpyAsync(s->d_params, s->h_stage, sizeof(double) * NNGP_NPARAM * K, cudaMemcpyHostToDevice, s->stream);
            if (e != cudaSuccess) return cuda_fail(s, e, "cudaMemcpyAsync");
            d_params = s->d_params;
        }
        PeerExchange px = s->px;
        px.gen = s->px.gen + 1;
        rc2 = launch_eval(s, kernel_id, d_params, params, K, nullptr, r == 0 ? s->h_out : nullptr, seq, world > 1 ? &px : nullptr, s->stream);
        if (!rc2) s->px.gen = px.gen;
        return rc2;
    });
    if (rc) return rc;
    rc = wait_lines(s0, s0->stream, K, seq, out);
    if (rc) h->err = s0->err;
    return rc;
}

int nngp_loglik(nngp_handle *h, int kernel_id, const double *params, int K, double *out)
{
    if (is_group(h)) return group_loglik(h, kernel_id, params, K, out);
    return loglik_host(h, kernel_id, params, K, out, false);
}

int nngp_loglik_allreduce(nngp_handle *h, int kernel_id, const double *params, int K, double *out)
{
    if (is_group(h)) return group_unsupported(h, "nngp_loglik_allreduce (nngp_loglik already returns the total over its devices)");
    return loglik_host(h, kernel_id, params, K, out, true);
}

int nngp_loglik_terms(nngp_handle *h, int kernel_id, double sigma2, double phi, double tau2, double *out3)
{
    const double prm[NNGP_NPARAM] = {sigma2, phi, tau2, 0.0};
    if (is_group(h)) return group_loglik(h, kernel_id, prm, 1, out3);
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    return loglik_host(h, kernel_id, prm, 1, out3, h->px.world > 1);
}

// shared by nngp_factors / nngp_cov_blocks: run the emitting variant over [i0, i1) in slabs
static int run_emit(nngp_handle *h, int kernel_id, const double *params, int64_t i0, int64_t i1, double *B,
                    double *F, double *CN, double *cc, double *cs)
{
    if (is_group(h)) {
        // rows are evaluated by the devices that hold them
        if (!h->has_nbr) return fail(h, NNGP_ESTATE, "no neighbour table");
        if (i0 < 0 || i1 < i0 || i1 > h->n) return fail(h, NNGP_EINVAL, "need 0 <= i0 <= i1 <= n");
        if (i0 < h->lo || i1 > h->hi) return fail(h, NNGP_ESTATE, "rows outside the shard of a multi-device handle have no neighbour sets");
        const int m = h->m;
        return group_each(h, [&](nngp_handle *s, int) {
            const int64_t a = std::max(i0, s->nbr_row0), b = std::min(i1, s->nbr_row0 + s->nbr_rows);
            if (a >= b) return int(NNGP_OK);
            const int64_t o = a - i0;
            return run_emit(s, kernel_id, params, a, b, B ? B + o * m : nullptr, F ? F + o : nullptr,
                            CN ? CN + o * m * m : nullptr, cc ? cc + o * m : nullptr, cs ? cs + o : nullptr);
        });
    }
    int rc = check_eval(h, kernel_id, params, 1);
    if (rc) return rc;
    if (i0 < 0 || i1 < i0 || i1 > h->n) return fail(h, NNGP_EINVAL, "need 0 <= i0 <= i1 <= n");
    if (i1 > i0 && !h->holds_rows(i0, i1)) return fail(h, NNGP_ESTATE, "the table of this handle holds the rows of its shard only");
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
    if (B) EMIT_TRY(dev_malloc_on(h, &dB, sizeof(double) * cap * m));
    if (F) EMIT_TRY(dev_malloc_on(h, &dF, sizeof(double) * cap));
    if (CN) EMIT_TRY(dev_malloc_on(h, &dCN, sizeof(double) * cap * m * m));
    if (cc) EMIT_TRY(dev_malloc_on(h, &dcc, sizeof(double) * cap * m));
    if (cs) EMIT_TRY(dev_malloc_on(h, &dcs, sizeof(double) * cap));
    if ((rc = ensure_scratch(h, 1, grid_for(h, kernel_id, cap), h->stream))) { cleanup(); return rc; }
    for (int64_t s0 = i0; s0 < i1; s0 += cap) {
        const int64_t s1 = s0 + cap < i1 ? s0 + cap : i1;
        const int64_t cnt = s1 - s0;
        EvalArgs a{};
        a.pts = h->pts; a.eps2 = h->eps2; a.nbr = h->nbr_base();
        a.lo = s0; a.hi = s1; a.m = m;
        a.params = nullptr; memcpy(a.pv, params, sizeof(double) * NNGP_NPARAM);
        a.partials = h->d_partials; a.counters = h->d_counters; a.out = h->d_out;
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
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!params) return fail(h, NNGP_EINVAL, "params must not be NULL");
    return run_emit(h, kernel_id, params, i0, i1, B, F, nullptr, nullptr, nullptr);
}

int nngp_cov_blocks(nngp_handle *h, int kernel_id, const double *params, int64_t i0, int64_t i1, double *CN,
                    double *cc, double *cs)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (!params) return fail(h, NNGP_EINVAL, "params must not be NULL");
    return run_emit(h, kernel_id, params, i0, i1, nullptr, nullptr, CN, cc, cs);
}

int nngp_set_timing(nngp_handle *h, int on)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (is_group(h)) {
        for (nngp_handle *s : h->group->subs) {
            const int rc = nngp_set_timing(s, on);
            if (rc) { h->err = s->err; return rc; }
        }
        return NNGP_OK;
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (on && !h->ev0) {
        CUDA_TRY(h, cudaEventCreate(&h->ev0));
        CUDA_TRY(h, cudaEventCreate(&h->ev1));
    }
    h->timing = on != 0;
    return NNGP_OK;
}

int nngp_last_eval_ms(nngp_handle *h, double *ms)
{
    if (!h || !ms) return fail(h, NNGP_EINVAL, "null argument");
    if (is_group(h)) {
        double worst = 0.0;
        for (nngp_handle *s : h->group->subs) {
            if (s->hi == s->lo) continue;
            double v = 0.0;
            const int rc = nngp_last_eval_ms(s, &v);
            if (rc) { h->err = s->err; return rc; }
            worst = std::max(worst, v);
        }
        *ms = worst;
        return NNGP_OK;
    }
    if (!h->timing || !h->ev1) return fail(h, NNGP_ESTATE, "timing is off (nngp_set_timing)");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaEventSynchronize(h->ev1));
    float f = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&f, h->ev0, h->ev1));
    *ms = double(f);
    return NNGP_OK;
}

int64_t nngp_launch_count(const nngp_handle *h)
{
    if (!h) return 0;
    if (h->group) {
        int64_t t = 0;
        for (const nngp_handle *s : h->group->subs) t += s->launches;
        return t;
    }
    return h->launches;
}

int nngp_measure_fma_peak(nngp_handle *h, int dtype, int iters, double *instr_per_s)
{
    if (!h) return fail(nullptr, NNGP_EINVAL, "null handle");
    if (is_group(h)) {
        const int rc = nngp_measure_fma_peak(h->group->subs[0], dtype, iters, instr_per_s);
        if (rc) h->err = h->group->subs[0]->err;
        return rc;
    }
    if (!instr_per_s || iters < 1) return fail(h, NNGP_EINVAL, "bad arguments");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, launch_fma_peak(h, dtype, iters, instr_per_s));
    return NNGP_OK;
}

}  // extern "C"
