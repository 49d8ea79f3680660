// loglik_fused.cuh -- stages 2+3 of the NNGP hot path as ONE kernel (sm_100a).
//
// Takes over the reference's per-location accessors (all stubs upstream):
//   _CNs nngp.py:78-82, _Ccross nngp.py:84-86, _Cs nngp.py:92-96  -> covariance build
//   _Bsi nngp.py:73-76, _Fsi nngp.py:88-90                         -> factorisation
//   and the reduction BASELINE.json north_star defines: sum_i [log F_i + r_i^2 / F_i].
//
// Mapping.  A location i with p <= m neighbours is the P x P augmented SPD matrix
//     M = [[C_N(i), c_i], [c_i^T, C(i,i)]],   rows 0..p-1 = neighbours (table order), rows p..P-2 =
//     identity padding, row P-1 = the location itself,
// with the right-hand side w = [y_N(i); 0; y_i].  An LDL^T elimination of M gives, at the last
// pivot, D_{P-1} = F_i and the eliminated w_{P-1} = r_i = y_i - b_i^T y_N(i): the log-likelihood
// needs no back-substitution.  G lanes share one location (P = G*R rows, each lane owning R of them -- folded, see
// row_of -- kept in registers); a warp carries 32/G locations.  Pivot columns travel through two alternating
// shared-memory buffers (one __syncwarp per pivot), everything else is lane-local FP64 (or FP32) arithmetic.
// Neighbour coordinates are gathered once per location as 32-byte records (cp.async, one iteration ahead) and
// staged in shared memory for the column reads of the covariance build.  Block partials are written per block and
// summed in a fixed order by the last block to finish (deterministic for a given grid).
//
// Bound (ncu): the unit closest to its peak is the shared-memory DATA STAGE (one 128-byte wavefront
// per cycle per SM; 84 % before round 2's layout work, 73 % after), then the FP64 pipe (51 %) and the issue slots
// (51 %); HBM traffic is the 4m+32 B/location of compulsory reads (3 % of peak).  Every shared-memory access here is
// laid out for its minimum wavefront count (WarpSmem).  Tensor cores do not apply: per-location work is a 16x16 /
// 32x32 factorisation with a serial pivot chain plus transcendental covariance entries.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "nngp_common.cuh"
#include "peer_exchange.cuh"

namespace nngp_fused {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;

#ifdef NNGP_TIMELINE
static __device__ unsigned long long nngp_tl[64];
static __device__ unsigned long long nngp_tl_blk[3][1024];  // per block: entry, loop done, smid
__device__ __forceinline__ void tl_stamp(int slot, bool who)
{
    if (who) {
        unsigned long long g;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        nngp_tl[slot] = g;
        nngp_tl[32 + slot] = clock64();
    }
}
#define TL(slot, who) tl_stamp(slot, who)
#else
#define TL(slot, who)
#endif

template <typename T, bool DIM3>
struct StagePt;
template <typename T>
struct __align__(2 * sizeof(T)) StagePt<T, false> {
    T x, y;
};
template <typename T>
struct __align__(16) StagePt<T, true> {
    T x, y, z, pad;
};

__device__ __forceinline__ double t_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float t_fma(float a, float b, float c) { return fmaf(a, b, c); }

// ---- branch-free math for the covariance build --------------------------------------------------
// The CUDA library sqrt()/exp()/division carry slow-path branches, 64-bit immediates and selects
// that cost more issue slots than FP64-pipe slots; the arguments here are known to be positive,
// finite and (for exp) non-positive, so the kernels below keep only the fast paths.

// exp(t) = 2^k * 2^(j/2^TB) * e^r, |r| <= ln2/2^(TB+1): a table of 2^TB entries in shared memory and a
// polynomial whose degree shrinks as the table grows (truncation < 4e-17 in all three):
//   TB = 11 (16 KB) cubic    -- (round 1; a random 8-byte lookup of a 2048-entry table costs 5.2 shared-memory
//                               wavefronts per warp -- 16 lanes of a half-warp into 16 bank pairs -- and the kernel is
//                               bound by the shared-memory data stage: profiles/r2a_fused_smem.txt)
//   TB =  8 quartic          -- the unrolled shapes: 16 COPIES of the 256 entries, copy c at [j * 16 + c], lane l
//                               reads copy l % 16 (REP = 4, 32 KB).  Every lane of a half-warp owns one bank pair:
//                               a lookup is conflict-free by construction, 2 wavefronts, for one more FMA per pair.
//                               The rolled (8, 4) shape keeps a single copy (2 KB): 2 blocks/SM must still fit
//   TB =  6 (512 B) quintic  -- the rolled (16, 3) shape
template <int TB> __host__ __device__ constexpr int exp_slot() { return TB == 11 ? 2 : TB == 8 ? 1 : 0; }
__constant__ double kExpA[3] = {-64.0 / 0.6931471805599453094, -256.0 / 0.6931471805599453094,
                                -2048.0 / 0.6931471805599453094};   // -2^TB / ln2
// -(ln2 / 2^TB): one FMA computes k*c - u with a single rounding; the constant's own rounding error
// (half an ulp) times k <= 2^TB u / ln2 stays below 1e-16 u, so no low-part correction is needed
__constant__ double kExpB[3] = {-0.6931471805599453094 / 64.0, -0.6931471805599453094 / 256.0,
                                -0.6931471805599453094 / 2048.0};

// 64-bit constants of the covariance math.  Read as constant-bank operands (c[3][..]) they cost no
// instruction; as literals ptxas re-materialises them (2 MOV each) inside the rolled build loop.
__constant__ double kCovC[12] = {
    0.375,                     // 0  sqrt correction
    0.0,                       // 1  (unused)
    6755399441055744.0,        // 2  1.5 * 2^52 (round-to-nearest magic)
    0.0,                       // 3  (unused)
    0.0,                       // 4  (unused)
    1.0 / 120.0,               // 5
    1.0 / 24.0,                // 6
    1.0 / 6.0,                 // 7
    1.0 / 3.0,                 // 8  Matern 5/2
    1e-290,                    // 9  tiny seed of the squared distance
    0.0, 0.0};

// u = sqrt(d2), d2 > 0 (the caller seeds the accumulation with a tiny positive constant).
// One MUFU.RSQ64H seed and a third-order correction: rel. error ~ 2 ulp.
__device__ __forceinline__ double fast_sqrt(double d2)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d2));
    const double t1 = d2 * y0;               // ~ sqrt(d2)
    const double e = fma(-t1, y0, 1.0);      // 1 - d2*y0^2
    const double c = fma(e, 0.375, 0.5) * e; // e/2 + 3e^2/8
    return fma(t1, c, t1);
}
__device__ __forceinline__ float fast_sqrt(float d2)
{
    float y0;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(d2));
    const float t1 = d2 * y0;
    const float e = fmaf(-t1, y0, 1.0f);
    return fmaf(t1 * 0.5f, e, t1);
}

// 1/x for a positive finite pivot: MUFU.RCP64H seed + one cubic correction (3 dependent FP64
// operations: this sits on the pivot-to-pivot critical path of the elimination)
__device__ __forceinline__ double fast_rcp(double x)
{
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    const double e = fma(-x, r0, 1.0);   // |e| <= 2^-20 (the seed reads the high word only)
    return fma(r0, fma(e, e, e), r0);    // r0 (1 + e + e^2): remainder e^3 < 2^-60
}
__device__ __forceinline__ float fast_rcp(float x)
{
    float r0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(x));
    const float e = fmaf(-x, r0, 1.0f);
    return fmaf(r0, e, r0);
}

// sigma2 * exp(-u) for u >= 0.  fp64: table of sigma2 * 2^(j/2^TB) in shared memory (tab) + a short
// polynomial -> 8 (TB = 8) to 9 (TB = 6) FP64 instructions.  u >= 708 (which includes the far-away
// sentinel rows) returns exactly 0 (this single-value form; the batch form caps u at 700 instead, see corr_batch).
template <int TB, int REP = 0>
__device__ __forceinline__ double scaled_exp_neg(double u, const double *tab, double /*sigma2*/)
{
    const double kd = fma(u, kExpA[exp_slot<TB>()], kCovC[2]);
    const int ki = __double2loint(kd);
    const double kf = kd - kCovC[2];                      // = round(-u * 2^TB/ln2)
    const double r = fma(kf, kExpB[exp_slot<TB>()], -u);
    double q;
    if (TB == 11) q = fma(r, 1.0 / 6.0, 0.5);
    else {
        q = TB == 8 ? fma(r, 1.0 / 24.0, 1.0 / 6.0) : fma(r, fma(r, 1.0 / 120.0, 1.0 / 24.0), 1.0 / 6.0);
        q = fma(r, q, 0.5);
    }
    q = fma(r, q, 1.0);
    const double tv = tab[(ki & ((1 << TB) - 1)) << REP];  // REP > 0: tab already points at this lane's copy
    const double p = fma(tv * r, q, tv);                  // tv * (1 + r*q)
    const int hi = __double2hiint(p) + ((ki >> TB) << 20); // * 2^k
    const double v = __hiloint2double(hi, __double2loint(p));
    return __double2hiint(u) >= 0x40862000 ? 0.0 : v;
}
template <int TB, int REP = 0>
__device__ __forceinline__ float scaled_exp_neg(float u, const float *, float sigma2)
{
    // k = round(-u / ln2) by the magic-number add (1.5 * 2^23): the sum's low mantissa bits are k itself, so
    // neither the rounding nor the float -> int conversion goes to the XU pipe, which the two MUFUs of a
    // pair (rsqrt, ex2) already load (ncu: XU 58 % active with FRND + F2I here, the busiest pipe of the kernel)
    const float t = fmaf(u, -1.4426950408889634f, 12582912.0f);
    const float kf = t - 12582912.0f;
    float r = fmaf(kf, -0.693145751953125f, -u);
    r = fmaf(kf, -1.428606765330187e-06f, r);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(r * 1.4426950408889634f));
    const float v = sigma2 * e * __int_as_float((__float_as_int(t) << 23) + 0x3f800000);  // * 2^k, |k| <= 126 below
    return u >= 87.0f ? 0.0f : v;
}

// sigma2 * rho(u), u = phi * distance (oracle: nngp_oracle.c corr()).
template <typename T, int KERN, int TB, int REP = 0>
__device__ __forceinline__ T cov_from_u(T u, const T *tab, T sigma2)
{
    const T e = scaled_exp_neg<TB, REP>(u, tab, sigma2);
    if (KERN == NNGP_EXPONENTIAL) return e;
    if (KERN == NNGP_MATERN32) return t_fma(u, e, e);
    return t_fma(u, t_fma(u, T(1.0 / 3.0), T(1)), T(1)) * e;
}

// The same arithmetic for a batch of B independent pairs, written stage by stage so that the
// instruction stream interleaves B dependency chains (the per-pair chain is ~20 dependent FP64
// operations; with 3 resident warps per scheduler a single chain leaves the FP64 pipe idle).
// In: x[b] = u_b^2 > 0.  Out: x[b] = sigma2 * rho(u_b).
// x[b] = u_b^2 > 0  ->  x[b] = u_b (MUFU.RSQ64H seed + third-order correction), B chains interleaved
template <int B>
__device__ __forceinline__ void sqrt_batch(double (&x)[B])
{
    double y0[B], t1[B], e[B];
#pragma unroll
    for (int b = 0; b < B; ++b) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0[b]) : "d"(x[b]));
#pragma unroll
    for (int b = 0; b < B; ++b) t1[b] = x[b] * y0[b];
#pragma unroll
    for (int b = 0; b < B; ++b) e[b] = fma(-t1[b], y0[b], 1.0);
#pragma unroll
    for (int b = 0; b < B; ++b) y0[b] = fma(e[b], kCovC[0], 0.5);
#pragma unroll
    for (int b = 0; b < B; ++b) y0[b] = y0[b] * e[b];
#pragma unroll
    for (int b = 0; b < B; ++b) x[b] = fma(t1[b], y0[b], t1[b]);
}
// x[b] = u_b >= 0  ->  x[b] = (table scale) * rho(u_b)
template <int KERN, int B, int TB, int REP = 0>
__device__ __forceinline__ void corr_batch(double (&x)[B], const double *tab)
{
    double e[B], kd[B], r[B], qq[B], tv[B];
    int ki[B];
    const double shift = kCovC[2];  // second constant of a two-constant FMA: a register
    // u is capped at 700 through its high word (one integer min; u >= 0): the exponent arithmetic below then never
    // leaves the normal range -- far-away sentinel rows included -- and e^-u bottoms out at 1e-304 instead of 0,
    // which no covariance can tell from 0.  (A compare and two selects on the result cost two more issue slots per pair.)
#pragma unroll
    for (int b = 0; b < B; ++b) x[b] = __hiloint2double(min(__double2hiint(x[b]), 0x4085E000), __double2loint(x[b]));
#pragma unroll
    for (int b = 0; b < B; ++b) kd[b] = fma(x[b], kExpA[exp_slot<TB>()], shift);
#pragma unroll
    for (int b = 0; b < B; ++b) { ki[b] = __double2loint(kd[b]); kd[b] = kd[b] - shift; }
#pragma unroll
    for (int b = 0; b < B; ++b) tv[b] = tab[(ki[b] & ((1 << TB) - 1)) << REP];
#pragma unroll
    for (int b = 0; b < B; ++b) r[b] = fma(kd[b], kExpB[exp_slot<TB>()], -x[b]);
    if (TB == 11) {
#pragma unroll
        for (int b = 0; b < B; ++b) qq[b] = fma(r[b], kCovC[7], 0.5);
    } else {
        if (TB == 8) {
#pragma unroll
            for (int b = 0; b < B; ++b) qq[b] = fma(r[b], kCovC[6], kCovC[7]);
        } else {
            const double c5 = kCovC[5];
#pragma unroll
            for (int b = 0; b < B; ++b) qq[b] = fma(r[b], c5, kCovC[6]);
#pragma unroll
            for (int b = 0; b < B; ++b) qq[b] = fma(r[b], qq[b], kCovC[7]);
        }
#pragma unroll
        for (int b = 0; b < B; ++b) qq[b] = fma(r[b], qq[b], 0.5);
    }
#pragma unroll
    for (int b = 0; b < B; ++b) qq[b] = fma(r[b], qq[b], 1.0);
#pragma unroll
    for (int b = 0; b < B; ++b) r[b] = tv[b] * r[b];
#pragma unroll
    for (int b = 0; b < B; ++b) r[b] = fma(r[b], qq[b], tv[b]);
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const int hi = __double2hiint(r[b]) + ((ki[b] >> TB) << 20);
        const double v = __hiloint2double(hi, __double2loint(r[b]));
        e[b] = v;
    }
#pragma unroll
    for (int b = 0; b < B; ++b) {
        if (KERN == NNGP_EXPONENTIAL) x[b] = e[b];
        else if (KERN == NNGP_MATERN32) x[b] = fma(x[b], e[b], e[b]);
        else x[b] = fma(x[b], fma(x[b], kCovC[8], 1.0), 1.0) * e[b];
    }
}
template <int KERN, int B, int TB, int REP = 0>
__device__ __forceinline__ void cov_batch(double (&x)[B], const double *tab, double)
{
    sqrt_batch<B>(x);
    corr_batch<KERN, B, TB, REP>(x, tab);
}
template <int KERN, int B, int TB, int REP = 0>
__device__ __forceinline__ void cov_batch(float (&x)[B], const float *tab, float sigma2)
{
#pragma unroll
    for (int b = 0; b < B; ++b) x[b] = cov_from_u<float, KERN, TB>(fast_sqrt(x[b]), tab, sigma2);
}

// far-away sentinel for padded rows: every covariance with it underflows (fp32: to exactly 0; fp64: to ~1e-304, the
// floor of the batch exponential, which no pivot or sum can tell from 0)
template <typename T> __device__ __forceinline__ T sentinel_coord(int r);
template <> __device__ __forceinline__ double sentinel_coord<double>(int r) { return 1e100 * double(r + 1); }
template <> __device__ __forceinline__ float sentinel_coord<float>(int r) { return 1e15f * float(r + 1); }
template <typename T> __device__ __forceinline__ T tiny_seed();
template <> __device__ __forceinline__ double tiny_seed<double>() { return kCovC[9]; }
template <> __device__ __forceinline__ float tiny_seed<float>() { return 1e-36f; }

__device__ __forceinline__ uint32_t smem_addr(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

template <typename T, int G>
__device__ __forceinline__ T grp_bcast(T v, int src)
{
    return __shfl_sync(0xffffffffu, v, src, G);
}

// Shared-memory tile of one location: the strict lower triangle of the augmented matrix, row-major
// (entry (a, b), a > b, at a(a-1)/2 + b).  The per-location stride is padded so the W locations of a
// warp start 16 bytes (mod 128) apart: no systematic bank aliasing between the lane groups.
// Column-major enumeration of the (row block s, column j) pairs a lane owns in the row-owner layout:
// column j is needed by the row blocks s >= (j + 1) / G.  Evaluated at compile time (unrolled callers).
template <int G, int R>
__host__ __device__ constexpr int pair_col(int t)
{
    int j = 0;
    for (; j < G * R - 1; ++j) {
        const int cnt = R - (j + 1) / G;
        if (t < cnt) break;
        t -= cnt;
    }
    return j;
}
template <int G, int R>
__host__ __device__ constexpr int pair_slot(int t)
{
    int j = 0;
    for (; j < G * R - 1; ++j) {
        const int cnt = R - (j + 1) / G;
        if (t < cnt) break;
        t -= cnt;
    }
    return (j + 1) / G + t;
}

// FOLD: the folded row assignment.  Lane q of a group owns row s*G + q in even row blocks and row s*G + (G-1-q) in
// odd ones.  A row needs the columns below it: every column of the blocks before its own (the same for all lanes)
// plus, inside its own (diagonal) block, q columns in an even block and G-1-q in an odd one -- G-1 per PAIR of
// blocks, whatever the lane.  So the diagonal blocks of a block pair (s0, s1) cost G-1 evaluation slots instead of
// 2(G-1): in slot t the lanes with q > t evaluate (row s0*G+q, column s0*G+t) and the others (row s1*G+G-1-q,
// column s1*G+G-2-t); both destinations are compile-time registers, and each is a don't-care entry (column >= row)
// for the lanes that did not mean it, so the value is stored to both without a select.  Every lane then evaluates
// exactly P(P-1)/2 / G pairs: (4,4) 30 instead of 36, (8,4) 62 instead of 76, (8,2) 15, (16,2) 31.
template <int G, int FOLD>
__host__ __device__ constexpr int row_of(int s, int q) { return (FOLD && (s & 1)) ? s * G + (G - 1 - q) : s * G + q; }
// full-block pairs of the folded layout, column-major: column j < G(R-1) is needed by the row blocks s > j / G
template <int G, int R>
__host__ __device__ constexpr int fpair_col(int t)
{
    int j = 0;
    for (; j < G * (R - 1); ++j) {
        const int cnt = R - 1 - j / G;
        if (t < cnt) break;
        t -= cnt;
    }
    return j;
}
template <int G, int R>
__host__ __device__ constexpr int fpair_slot(int t)
{
    int j = 0;
    for (; j < G * (R - 1); ++j) {
        const int cnt = R - 1 - j / G;
        if (t < cnt) break;
        t -= cnt;
    }
    return j / G + 1 + t;
}
// one evaluation slot of the unrolled build: a full-block pair (row block s, column j), or -- FOLD only -- slot tt of
// the diagonal blocks of the block pair (s, s1)
struct SlotInfo {
    bool diag;
    int s, j;    // full: row block, column.  diag: first block of the pair, its column s*G + tt
    int s1, j1;  // diag: second block of the pair, its column s1*G + G-2-tt
    int tt;
};
template <int G, int R, int FOLD>
__host__ __device__ constexpr int n_slots() { return FOLD ? G * R * (R - 1) / 2 + (R / 2) * (G - 1) : G * R * (R + 1) / 2 - R; }
// Slot order of the folded layout.  FOLD = 1, column-major over the row blocks: consecutive slots share a column, whose
// staged coordinates are then loaded once for up to R-1 slots.  FOLD = 2, row-block-major: block after block, the
// diagonal slots of a block pair right after its second block -- the scaled coordinates of a row block are dead once
// its slots are done, while the matrix registers fill up, which lowers the peak register need; every slot loads its
// own column (cheap for 16-byte 2-D points, not for 32-byte 3-D ones).
template <int G, int R, int FOLD>
__host__ __device__ constexpr SlotInfo slot_info(int t)
{
    if (!FOLD) return SlotInfo{false, pair_slot<G, R>(t), pair_col<G, R>(t), 0, 0, 0};
    if (FOLD == 2) {
        for (int s = 1; s < R; ++s) {
            if (t < s * G) return SlotInfo{false, s, t, 0, 0, 0};
            t -= s * G;
            if (s & 1) {
                if (t < G - 1) return SlotInfo{true, s - 1, (s - 1) * G + t, s, s * G + G - 2 - t, t};
                t -= G - 1;
            }
        }
        return SlotInfo{false, R - 1, 0, 0, 0, 0};  // not reached for t < n_slots
    }
    constexpr int NFULL = G * R * (R - 1) / 2;
    if (t < NFULL) return SlotInfo{false, fpair_slot<G, R>(t), fpair_col<G, R>(t), 0, 0, 0};
    const int u = t - NFULL, p = u / (G - 1), tt = u % (G - 1);
    return SlotInfo{true, 2 * p, 2 * p * G + tt, 2 * p + 1, (2 * p + 1) * G + G - 2 - tt, tt};
}

// BUILD = 6: the sweep variant of the unrolled build.  A block evaluates a chunk of up to kSweepChunk
// parameter vectors per location: the pair distances (d^2 and sqrt: 9 of the 17 FP64 instructions a
// covariance entry costs) are computed once and kept in registers, and for every parameter vector only
// the correlation function, the elimination and the accumulation run.  sigma2 is factored out
// (C = sigma2 (R + delta I), delta_i = (tau2 + eps2_i) / sigma2): the exp table carries no parameter, and
// sum log F = (n - n_bad) log sigma2 + sum log F', sum r^2/F = (sum r^2/F') / sigma2 are restored at the end.
constexpr int kSweepChunk = 8;
template <int BUILD>
__host__ __device__ constexpr bool sweep_build() { return BUILD == 6; }

template <int P>
__host__ __device__ constexpr int tile_stride() { return (P * (P - 1) / 2 + 7) / 8 * 8 + 4; }

// Per-warp shared-memory carve-up (bytes).  Everything a warp touches is private to it, so the
// main loop needs __syncwarp only.
template <typename T, int G, int R, bool DIM3, int BUILD, int ELIM = 1>
struct WarpSmem {
    static constexpr int P = G * R, W = 32 / G;
    // location strides are padded by 16 bytes so the W groups of a warp start in different banks
    // (unpadded, all groups alias: 8-way conflicts on every record / coordinate access)
    static constexpr int rec_stride = P * 32 + 16;                        // bytes per location
    // Staged (re-centred, scaled) coordinates.  fp64 in 3-D: {x, y} pairs and z in separate arrays (a 32-byte
    // {x, y, z, pad} point makes every 16-byte store a 2-way bank conflict and every column read two 16-byte loads).
    static constexpr bool SOA3 = DIM3 && sizeof(T) == 8;
    static constexpr int pt_bytes = SOA3 ? 16 : int(sizeof(StagePt<T, DIM3>));
    static constexpr int stage_stride = P * pt_bytes + 16;
    // elements per location of the z array (SOA3) and of the elimination's column buffers: P entries, w_k, pad
    // (fp32, G = 4: 4-byte entries; a stride of P + 4 makes the eight 4-row windows of a warp tile the banks -- P + 2
    // gave every store two wavefronts)
    static constexpr int col_stride = P + ((sizeof(T) == 4 && G == 4) ? 4 : 2);
    static constexpr size_t zbuf = SOA3 ? size_t(sizeof(T) == 8 && G == 8 ? 6 : W) * col_stride * sizeof(T) : 0;
    // Where location g of the warp sits inside those arrays.  The kernel is bound by the shared-memory data stage
    // (profiles/r2a_fused_smem.txt), and with the W = 8 locations of a (4, R) shape in plain order the row-owner
    // stores collide: an 8-byte store is processed half a warp at a time and its 4 locations' windows of 4 consecutive
    // rows must tile the 32 banks, while an 8-byte broadcast load sees all 8 locations at once; a 16-byte store is
    // processed a quarter-warp (2 locations) at a time, a 16-byte broadcast load half a warp.  With strides of 16 bytes
    // (mod 128) the slot ORDER decides which locations meet: measured (tools/ubench/smem_layout.cu) 2.0 / 1.3 / 2.0
    // wavefronts per column store / pivot load / column load with col_pos against 4.0 / 1.3 / 2.0 in plain order,
    // and 4.6 / 2.0 per staged-point store / load with stage_pos against 8.0 / 2.0 -- the minima of these accesses.
    // (G = 8, W = 4: the two locations of a half-warp 64 bytes apart (mod 128) for the column stores -- slots 0, 4, 1, 5
    // of 6; its staged points are fine in plain order.)
    static constexpr bool PERM = sizeof(T) == 8 && G == 4, PERM8 = sizeof(T) == 8 && G == 8;
    static constexpr int col_slots = PERM8 ? 6 : W;
    __host__ __device__ static constexpr int col_pos(int g) { return PERM ? (g & 3) * 2 + (g >> 2) : PERM8 ? (g & 1) * 4 + (g >> 1) : g; }
    // fp32 with G = 4: its 2-D staged points are 8-byte entries at a 16-byte (mod 128) stride -- the case of the fp64
    // column buffers -- and its 3-D ones 16-byte entries like the fp64 2-D points
    static constexpr bool PERM32 = sizeof(T) == 4 && G == 4;
    __host__ __device__ static constexpr int stage_pos(int g)
    {
        return (PERM || (PERM32 && DIM3)) ? (g & 1) * 4 + (g >> 1) : (PERM32 && !DIM3) ? (g & 3) * 2 + (g >> 2) : g;
    }
    static constexpr size_t rec = size_t(W) * rec_stride;                 // gathered {x,y,z,yval} records
    static constexpr size_t e2 = DIM3 ? size_t(W) * P * sizeof(double) : 0;  // gathered eps2 (D = 3 only;
                                                                             // D < 3 records carry it in .z)
    static constexpr size_t idx = size_t(R) * 32 * sizeof(int);          // next group's neighbour indices
    // per-lane partial sums: {mantissa product, sum r^2/F, exponent sum} (+ n_bad and one set per
    // parameter vector of the chunk in the sweep variant)
    static constexpr size_t acc = sweep_build<BUILD>() ? size_t(kSweepChunk) * 4 * 32 * sizeof(double) : size_t(3) * 32 * sizeof(double);
    static constexpr size_t tile = BUILD == 1 ? (size_t(W) * tile_stride<P>() * sizeof(T) + 15) / 16 * 16 : 0;  // pair-indexed tile
    // scaled coordinates during the build; the elimination's two column buffers afterwards
    static constexpr size_t colbuf = size_t(2) * col_slots * col_stride * sizeof(T);
    static constexpr size_t stage = (size_t(W) * stage_stride + zbuf > colbuf ? size_t(W) * stage_stride + zbuf : colbuf);
    static constexpr size_t dump = size_t(P) * P * sizeof(T);            // emit only: one location's factor
    __host__ __device__ static constexpr size_t total(bool emit) { return rec + e2 + idx + acc + tile + stage + (emit ? dump : 0); }
};

// block-shared part: exp table + the launch's pair list (one packed word per pair)
template <int G, int BUILD>
__host__ __device__ constexpr int exp_tab_bits() { return BUILD != 1 ? 8 : (G == 16 ? 6 : 8); }
// log2 of the copies of the table (one per lane of a half-warp in the unrolled shapes)
template <int G, int BUILD>
__host__ __device__ constexpr int exp_rep_bits() { return BUILD != 1 ? 4 : 0; }
// pairs evaluated in lock step by one lane: 8 in the rolled (8, 4) build, whose loop is bound by the
// latency of its dependent shared-memory reads (pair list -> coordinates -> exp table)
template <int G, int BUILD>
__host__ __device__ constexpr int cov_batch_size() { return BUILD == 1 && G == 8 ? 8 : 4; }
template <typename T, int G, int BUILD>
__host__ __device__ constexpr size_t exp_tab_bytes() { return sizeof(T) == 8 ? (sizeof(double) << (exp_tab_bits<G, BUILD>() + exp_rep_bits<G, BUILD>())) : 16; }  // fp32: no table
template <typename T, int G, int R, int BUILD>
__host__ __device__ constexpr size_t block_smem()
{
    return exp_tab_bytes<T, G, BUILD>() + (BUILD == 1 ? (size_t(G * R) * (G * R - 1) / 2 + G * cov_batch_size<G, BUILD>()) * 8 : 0);
}

// dynamic shared memory needed by one block
template <typename T, int G, int R, bool DIM3, int BUILD, int ELIM = 1>
constexpr size_t smem_bytes(bool emit)
{
    return block_smem<T, G, R, BUILD>() + size_t(kWarps) * WarpSmem<T, G, R, DIM3, BUILD, ELIM>::total(emit);
}

template <typename T> struct Pair2;
template <> struct Pair2<double> { using type = double2; };
template <> struct Pair2<float> { using type = float2; };

// ELIM = 0: pivot-column entries travel by width-G warp shuffles.  ELIM = 1: the lanes publish the
// column in shared memory once per pivot and read it back as broadcast 2-element loads -- a quarter
// of the instructions the 64-bit shuffles (two SHFL plus their register moves each) cost.
// EMIT: the variant that also writes per-location outputs (accessors, factors(), prediction); compiled
// separately so the metric's kernel carries none of its code or registers.
template <typename T, int G, int R, int KERN, bool DIM3, int MINB, int BUILD, int ELIM, bool EMIT, int FOLD>
__global__ void __launch_bounds__(kThreads, MINB) fused_loglik_kernel(const __grid_constant__ EvalArgs a)
{
    constexpr int P = G * R;   // rows of the augmented matrix
    constexpr int W = 32 / G;  // locations per warp
    static_assert(!FOLD || (R % 2 == 0 && ELIM == 1), "the folded layout pairs row blocks");
    using WS = WarpSmem<T, G, R, DIM3, BUILD, ELIM>;
    constexpr bool SOA3 = WS::SOA3;          // 3-D fp64: z staged in its own array
    using Pt = StagePt<T, DIM3 && !SOA3>;

    extern __shared__ __align__(32) unsigned char smem_raw[];
    TL(0, blockIdx.x == 0 && threadIdx.x == 0);
#ifdef NNGP_TIMELINE
    if (threadIdx.x == 0) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); nngp_tl_blk[0][blockIdx.x] = g_; unsigned int sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); nngp_tl_blk[2][blockIdx.x] = sm_; }
#endif
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int q = lane % G;  // this lane's position in its group: it owns rows row_of(s, q)
    const int g = lane / G;  // location slot inside the warp
    auto rowq = [&](int s) { return row_of<G, FOLD>(s, q); };

    T *exp_tab = reinterpret_cast<T *>(smem_raw);  // sigma2 * 2^(j/64) (fp64 path only)
    uint2 *pair_lut = reinterpret_cast<uint2 *>(smem_raw + exp_tab_bytes<T, G, BUILD>());
    unsigned char *wbase = smem_raw + block_smem<T, G, R, BUILD>() + size_t(warp) * WS::total(EMIT);
    unsigned char *recbuf = wbase + g * WS::rec_stride;  // this location's records (16-byte aligned)
    double *e2buf = reinterpret_cast<double *>(wbase + WS::rec);
    int *idxbuf = reinterpret_cast<int *>(wbase + WS::rec + WS::e2) + lane;          // [s * 32]
    volatile double *accbuf = reinterpret_cast<volatile double *>(wbase + WS::rec + WS::e2 + WS::idx) + lane;
    T *tile_w = reinterpret_cast<T *>(wbase + WS::rec + WS::e2 + WS::idx + WS::acc);
    T *tile = tile_w + g * tile_stride<P>();  // this location's covariance entries
    Pt *stage = reinterpret_cast<Pt *>(wbase + WS::rec + WS::e2 + WS::idx + WS::acc + WS::tile + WS::stage_pos(g) * WS::stage_stride);
    [[maybe_unused]] T *zstage = reinterpret_cast<T *>(wbase + WS::rec + WS::e2 + WS::idx + WS::acc + WS::tile + WS::W * WS::stage_stride) + WS::col_pos(g) * WS::col_stride;
    T *dump = reinterpret_cast<T *>(wbase + WS::rec + WS::e2 + WS::idx + WS::acc + WS::tile + WS::stage);
    // column buffers of the elimination: alias the staged coordinates (dead once the build is done)
    T *colw = reinterpret_cast<T *>(wbase + WS::rec + WS::e2 + WS::idx + WS::acc + WS::tile);
    T *col_even = colw + WS::col_pos(g) * WS::col_stride, *col_odd = colw + (WS::col_slots + WS::col_pos(g)) * WS::col_stride;

    constexpr bool SWEEP = sweep_build<BUILD>();
    static_assert(!SWEEP || (sizeof(T) == 8 && !EMIT && ELIM == 1), "the sweep variant is fp64, reduction only");
    // FACT: sigma2 is factored out of the matrix, C = sigma2 (R + delta I) with delta_i = (tau2 + eps2_i) / sigma2,
    // and restored in the last block (sum log F = (n - n_bad) log sigma2 + sum log F', sum r^2/F = (sum r^2/F') /
    // sigma2).  The exp table then carries no parameter: it is an asynchronous copy of the handle's 2^(j/2048)
    // table that overlaps the first gather instead of a load-multiply-store pass every block waits for.  The
    // emitting variant keeps sigma2 inside (its outputs are the covariances themselves), and so does fp32.
    constexpr bool FACT = sizeof(T) == 8 && !EMIT;
    // sweep: blockIdx.y = chunk of parameter vectors [k0, k0 + kc); otherwise one vector per blockIdx.y
    const int k0 = SWEEP ? int(blockIdx.y) * kSweepChunk : int(blockIdx.y);
    const int kc = SWEEP ? (a.K - k0 < kSweepChunk ? a.K - k0 : kSweepChunk) : 1;
    // parameter vector k, component c: from device memory, or straight from the kernel arguments (a host-pointer
    // call with K <= NNGP_PV_MAX needs no H2D copy: the values sit in the constant bank when the block starts)
    auto prm_at = [&](int k, int c) -> double { return a.params ? a.params[size_t(k) * NNGP_NPARAM + c] : a.pv[k][c]; };
    const T sigma2 = FACT ? T(1) : T(prm_at(k0, 0));
    const double phi = SWEEP ? 1.0 : prm_at(k0, 1);  // sweep: coordinates stay unscaled, u = phi_k * distance per vector
    // {diagonal without eps2, scale of eps2}: read from shared memory where they are used -- the main loop is at
    // its register limit and pays for every value kept live across it
    __shared__ double s_diag[2];
    if (threadIdx.x == 0) {
        const double p0 = prm_at(k0, 0), p2 = prm_at(k0, 2);
        const double inv_s2 = FACT && !SWEEP ? 1.0 / p0 : 1.0;
        s_diag[0] = FACT && !SWEEP ? fma(p2, inv_s2, 1.0) : p0 + p2;
        s_diag[1] = inv_s2;
    }
    const int m = a.m;
    constexpr int TB = exp_tab_bits<G, BUILD>(), REP = exp_rep_bits<G, BUILD>();
    // (the sweep variant keeps the synchronous fill: its launches are long, and its register allocation is too
    // finely balanced to touch -- the asynchronous form cost it 10 % in spills)
    constexpr bool ASYNC_TAB = FACT && REP == 4 && TB == 8 && !SWEEP;
    if constexpr (sizeof(T) == 8) {
        if constexpr (ASYNC_TAB) {
            // 32 KB (the handle's replicated table, a.exp2tab + 2048), 16 bytes per cp.async, L2-only (every block
            // reads the same lines); completes under the prologue's first wait_group, published to the block by
            // the __syncthreads below
            for (int k = threadIdx.x; k < (1 << (TB + REP)) / 2; k += kThreads)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(exp_tab + 2 * k)), "l"(a.exp2tab + 2048 + 2 * k) : "memory");
        } else {  // sigma2 * 2^(j / 2^TB) from the handle's table of 2^(j/2048): one L2 load and a multiply
            for (int k = threadIdx.x; k < (1 << (TB + REP)); k += kThreads)
                exp_tab[k] = T(double(sigma2) * __ldg(a.exp2tab + ((k >> REP) << (11 - TB))));
        }
    }
    const T *exp_lane = exp_tab + (REP ? (lane & ((1 << REP) - 1)) : 0);  // this lane's copy of the table
    // sum log F is carried as log(prod of mantissas) + ln2 * (sum of exponents): one multiply and a few
    // integer operations per location instead of a log() the whole warp would issue for one lane in G
    int nbad = 0;
    __shared__ double sw_prm[kSweepChunk][4];  // sweep: {phi, tau2 / sigma2, 1 / sigma2, log sigma2} per vector
    if constexpr (SWEEP) {
        if (threadIdx.x < kc) {
            const int kv = k0 + int(threadIdx.x);
            const double pk0 = prm_at(kv, 0);
            sw_prm[threadIdx.x][0] = prm_at(kv, 1);
            sw_prm[threadIdx.x][1] = prm_at(kv, 2) / pk0;
            sw_prm[threadIdx.x][2] = 1.0 / pk0;
            sw_prm[threadIdx.x][3] = log(pk0);
        }
        for (int kk = 0; kk < kSweepChunk; ++kk) {
            accbuf[kk * 128] = 1.0; accbuf[kk * 128 + 32] = 0.0; accbuf[kk * 128 + 64] = 0.0; accbuf[kk * 128 + 96] = 0.0;
        }
    } else {
        accbuf[0] = 1.0; accbuf[32] = 0.0; accbuf[64] = 0.0;  // mantissa product, sum r^2/F, exponent sum: rarely
                                                                // touched, kept out of the register file
    }
    // The launch's pair list: rows in play are the m neighbour rows and the location's own row P-1
    // (rows m .. P-2 are identity padding: their entries stay at the zero the tile is filled with).
    // Pair t -> packed {row a, column b, tile offset a(a-1)/2 + b}; lanes walk the list with stride G,
    // which splits the build evenly (m = 15: 30 pairs per lane) whatever the row-owner layout is.
    constexpr int CB = cov_batch_size<G, BUILD>();  // pairs evaluated in lock step by one lane
    const int rows_in_play = m + 1;
    const int npairs = rows_in_play * (rows_in_play - 1) / 2;
    const int nbatch = (npairs + G * CB - 1) / (G * CB);  // build-loop trips; the list is padded to it
    if (BUILD == 1) {
        for (int t = threadIdx.x; t < nbatch * G * CB; t += kThreads) {
            const int tt = t < npairs ? t : npairs - 1;  // padding repeats the last pair (same value, same slot)
            int ia = int((1.0f + sqrtf(1.0f + 8.0f * float(tt))) * 0.5f);
            while (ia * (ia - 1) / 2 > tt) --ia;
            while ((ia + 1) * ia / 2 <= tt) ++ia;
            const int ib = tt - ia * (ia - 1) / 2;
            const int ra = (ia == rows_in_play - 1) ? P - 1 : ia;
            // byte offsets: row point | column point << 16 inside stage[], entry inside the tile
            pair_lut[t] = make_uint2(uint32_t(ra * sizeof(Pt)) | (uint32_t(ib * sizeof(Pt)) << 16),
                                     uint32_t((ra * (ra - 1) / 2 + ib) * sizeof(T)));
        }
        for (int k = lane; k < W * tile_stride<P>(); k += 32) tile_w[k] = T(0);
    }
    __syncthreads();
    TL(1, blockIdx.x == 0 && threadIdx.x == 0);

    const int64_t nloc = a.hi - a.lo;
    const int64_t ngroups = (nloc + W - 1) / W;
    const int64_t gstride = int64_t(gridDim.x) * kWarps;
    const int64_t grp0 = int64_t(blockIdx.x) * kWarps + warp;

    // ---- software-pipelined gather, all through cp.async (LDGSTS) ---------------------------------
    // While group t is being built and eliminated, the 32-byte records of group t+1 and the
    // neighbour indices of group t+2 are in flight global -> shared.  Nothing loop-carried ever waits
    // on a global load: at the top of an iteration one cp.async.wait_group covers copies that were
    // issued a whole iteration earlier.
    auto issue_idx = [&](int64_t grp) {  // indices of `grp` -> idxbuf (own slots only)
        const int64_t i = a.lo + grp * W + g;
        const bool live = grp < ngroups && i < a.hi;
#pragma unroll
        for (int s = 0; s < R; ++s) {
            const int r = rowq(s);
            if (live && r < m && r != P - 1) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(idxbuf + s * 32)),
                             "l"(a.nbr + i * m + r)
                             : "memory");
            } else {
                idxbuf[s * 32] = (live && r == P - 1) ? int(i) : -1;
            }
        }
    };
    int nidx[R];
    auto issue_rec = [&]() {  // idxbuf -> registers; records of those rows -> recbuf
#pragma unroll
        for (int s = 0; s < R; ++s) nidx[s] = idxbuf[s * 32];
#pragma unroll
        for (int s = 0; s < R; ++s) {
            if (nidx[s] >= 0) {
                const int slot = g * P + rowq(s);
                const uint32_t dst = smem_addr(recbuf + rowq(s) * 32);
                const double4 *src = a.pts + nidx[s];
                // Records that fit the L2 are gathered through the L1 (.ca: the second half of a record hits the sector
                // its first half brought in).  When the records are several times the L2 most gathers come from HBM and
                // allocating their lines in the L1 only evicts what the block still needs: .cg then (measured: cfg3
                // 0.397 ms with .ca against 0.417 with .cg; cfg4 25.8 ms with .ca against 24.1 with .cg).
                if (a.gather_bypass_l1) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u),
                                 "l"(reinterpret_cast<const char *>(src) + 16)
                                 : "memory");
                } else {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16u),
                                 "l"(reinterpret_cast<const char *>(src) + 16)
                                 : "memory");
                }
                if (DIM3 && a.eps2)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(e2buf + slot)),
                                 "l"(a.eps2 + nidx[s])
                                 : "memory");
            }
        }
    };
    issue_idx(grp0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if constexpr (ASYNC_TAB) __syncthreads();  // every thread's slice of the exp table has landed
    issue_rec();
    issue_idx(grp0 + gstride);
    asm volatile("cp.async.commit_group;" ::: "memory");

    for (int64_t grp = grp0; grp < ngroups; grp += gstride) {
        const int64_t i = a.lo + grp * W + g;
        const bool live = i < a.hi;

        // ---- consume the prefetched records ------------------------------------------------------
        // Coordinates are re-centred on the location (row P-1) and scaled by phi in fp64 before any
        // cast, so neighbour differences keep full relative accuracy (also in fp32 mode) and the
        // covariance build works directly on u^2 = (phi*d)^2.
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        TL(2, blockIdx.x == 0 && threadIdx.x == 0 && grp == grp0);
        T rx[R], ry[R], rz[R], w[R], dg[R];
        T w0[R], e2r[R];  // sweep only: the untouched right-hand side and eps2 of the lane's rows
        bool valid[R];
        {
            const double2 *recs = reinterpret_cast<const double2 *>(recbuf);
            const double2 self0 = recs[2 * (P - 1)];
            const double sx = self0.x, sy = self0.y;
            const double sz = DIM3 ? recs[2 * (P - 1) + 1].x : 0.0;
            const double diag0 = SWEEP ? 1.0 : s_diag[0], inv_s2 = SWEEP ? 1.0 : s_diag[1];  // sweep: set per vector below
            bool mine = true;
#pragma unroll
            for (int s = 0; s < R; ++s) mine &= nidx[s] >= 0;
            const bool all_valid = __all_sync(0xffffffffu, mine);  // warp-uniform
#pragma unroll
            for (int s = 0; s < R; ++s) {
                const int r = rowq(s);
                valid[s] = nidx[s] >= 0;
                double2 v0 = make_double2(0.0, 0.0), v1 = make_double2(0.0, 0.0);
                double e2 = 0.0;
                if (all_valid) {
                    // every row of every location of the warp exists (any group past the first few rows of the ordering
                    // when m = P - 1): no padding, so no selects
                    v0 = recs[2 * r];
                    v1 = recs[2 * r + 1];
                    if (!DIM3) e2 = v1.x;
                    else if (a.eps2) e2 = e2buf[g * P + r];
                    rx[s] = T((v0.x - sx) * phi);
                    ry[s] = T((v0.y - sy) * phi);
                    rz[s] = T((v1.x - sz) * phi);
                    w[s] = T(v1.y);
                    dg[s] = T(FACT && !SWEEP ? fma(e2, inv_s2, diag0) : diag0 + e2);
                } else {
                if (valid[s]) {
                    v0 = recs[2 * r];
                    v1 = recs[2 * r + 1];
                    if (!DIM3) e2 = v1.x;  // D < 3: the record's z slot carries eps2
                    else if (a.eps2) e2 = e2buf[g * P + r];
                }
                // padded rows sit at a far-away sentinel: all their covariances vanish (see sentinel_coord)
                rx[s] = valid[s] ? T((v0.x - sx) * phi) : sentinel_coord<T>(r);
                ry[s] = valid[s] ? T((v0.y - sy) * phi) : T(0);
                rz[s] = valid[s] ? T((v1.x - sz) * phi) : T(0);
                w[s] = valid[s] ? T(v1.y) : T(0);
                dg[s] = valid[s] ? T(FACT && !SWEEP ? fma(e2, inv_s2, diag0) : diag0 + e2) : T(1);
                }
                w0[s] = w[s];
                e2r[s] = T(e2);
                Pt pt;
                pt.x = rx[s]; pt.y = ry[s];
                if constexpr (SOA3) zstage[r] = rz[s];
                else if constexpr (DIM3) { pt.z = rz[s]; pt.pad = T(0); }
                stage[r] = pt;
            }
        }
        __syncwarp();  // recbuf fully read, stage[] fully written
        issue_rec();                      // records of group t+1 (indices arrived an iteration ago)
        issue_idx(grp + 2 * gstride);     // indices of group t+2
        asm volatile("cp.async.commit_group;" ::: "memory");

        // ---- stage 2: covariance build ------------------------------------------------------------
        T A[R][P];
        // The diagonal entry of the lane's row in block s sits in column s*G + q (even blocks) or s*G + G-1-q (odd blocks
        // of the folded layout): a register array cannot be indexed by the lane, so it is placed by selects over the
        // block's columns.  The block's last column is a diagonal only for the lane whose row it is and a don't-care
        // entry for every other lane: no select there.
        auto set_diag = [&](T (&M)[R][P], const T (&dgv)[R]) {
#pragma unroll
            for (int s = 0; s < R; ++s) {
#pragma unroll
                for (int c = 0; c < G - 1; ++c) M[s][s * G + c] = (rowq(s) == s * G + c) ? dgv[s] : M[s][s * G + c];
                M[s][s * G + G - 1] = dgv[s];
            }
        };
        [[maybe_unused]] T Dst[SWEEP ? n_slots<G, R, FOLD>() : 1];  // sweep: the lane's pair distances
        if constexpr (BUILD == 1) {
        // A rolled loop over this lane's share of the pair list, CB pairs in lock step (independent
        // dependency chains: one chain cannot fill the FP64 pipe -- 8-cycle DFMA latency).  Operands
        // come from the staged coordinates, results go to the location's tile: nothing here needs a
        // register index, so the loop body stays small (instruction cache) and holds few registers.
            {
                // the list pointer is rebuilt from %laneid here on purpose: as an ordinary loop-carried value
                // ptxas spills it across the elimination and its local-memory reload stalled every
                // iteration (10 % of the kernel in the ncu source view)
                uint32_t lane_now;
                asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane_now));
                const uint2 *lp = pair_lut + (lane_now % G);
                const unsigned char *sb = reinterpret_cast<const unsigned char *>(stage);
                unsigned char *tb = reinterpret_cast<unsigned char *>(tile);
    #pragma unroll 1
                for (int u = 0; u < nbatch; ++u, lp += G * CB) {
                    T d2[CB];
                    uint32_t eoff[CB];
    #pragma unroll
                    for (int b = 0; b < CB; ++b) {
                        const uint2 pw = lp[b * G];
                        const Pt pa = *reinterpret_cast<const Pt *>(sb + (pw.x & 0xffffu));
                        const Pt pb = *reinterpret_cast<const Pt *>(sb + (pw.x >> 16));
                        eoff[b] = pw.y;
                        const T dx = pa.x - pb.x, dy = pa.y - pb.y;
                        d2[b] = t_fma(dy, dy, t_fma(dx, dx, tiny_seed<T>()));
                        if constexpr (SOA3) {  // z entries are half the size of the {x, y} pairs: half the byte offset
                            const unsigned char *zb = reinterpret_cast<const unsigned char *>(zstage);
                            const T dz = *reinterpret_cast<const T *>(zb + ((pw.x & 0xffffu) >> 1)) - *reinterpret_cast<const T *>(zb + (pw.x >> 17));
                            d2[b] = t_fma(dz, dz, d2[b]);
                        } else if constexpr (DIM3) { const T dz = pa.z - pb.z; d2[b] = t_fma(dz, dz, d2[b]); }
                    }
                    cov_batch<KERN, CB, TB, REP>(d2, exp_lane, sigma2);
    #pragma unroll
                    for (int b = 0; b < CB; ++b) *reinterpret_cast<T *>(tb + eoff[b]) = d2[b];
                }
            }
            __syncwarp();  // the tile is complete: lanes now read entries other lanes of the group built

            // the lane's rows come from the tile into registers for the elimination
    #pragma unroll
            for (int s = 0; s < R; ++s) {
                const int r = rowq(s);
                const T *row = tile + r * (r - 1) / 2;
    #pragma unroll
                for (int j = 0; j < P; ++j)
                    if (j < s * G + G) A[s][j] = j < r ? row[j] : T(0);  // entries right of the diagonal: don't care
            }
            set_diag(A, dg);
        } else {
            // Row-owner build, fully unrolled.  The lane's pairs (row block s, column j) are enumerated
            // column-major at compile time and evaluated CB at a time in lock step -- always CB
            // independent chains, whichever rows/columns they belong to (batching by column alone runs
            // out of parallelism on the late columns, which only the last row block needs) -- and the
            // results go straight to the row registers.
            // BUILD selects the unrolled build's batch width: 0 -> 4, 2 -> 6, 3 -> 9, 4 -> 12 pairs in lock step
            constexpr int CB = BUILD == 0 ? 4 : (BUILD == 2 || BUILD == 6) ? 6 : BUILD == 3 ? 9 : 12;
            constexpr int NP = n_slots<G, R, FOLD>();
#pragma unroll
            for (int t0 = 0; t0 < NP; t0 += CB) {
                T d2[CB];
#pragma unroll
                for (int b = 0; b < CB; ++b) {
                    const int t = (t0 + b < NP) ? t0 + b : NP - 1;
                    const SlotInfo si = slot_info<G, R, FOLD>(t);
                    T ax, ay, az;
                    Pt cj;
                    [[maybe_unused]] T cz = T(0);
                    if (si.diag) {
                        // lanes q > tt: (row of block s, column j); the others: (row of block s1, column j1)
                        const bool first = q > si.tt;
                        ax = first ? rx[si.s] : rx[si.s1];
                        ay = first ? ry[si.s] : ry[si.s1];
                        az = first ? rz[si.s] : rz[si.s1];
                        cj = stage[first ? si.j : si.j1];
                        if constexpr (SOA3) cz = zstage[first ? si.j : si.j1];
                    } else {
                        ax = rx[si.s]; ay = ry[si.s]; az = rz[si.s];
                        cj = stage[si.j];
                        if constexpr (SOA3) cz = zstage[si.j];
                    }
                    const T dx = ax - cj.x, dy = ay - cj.y;
                    d2[b] = t_fma(dy, dy, t_fma(dx, dx, tiny_seed<T>()));
                    if constexpr (SOA3) { const T dz = az - cz; d2[b] = t_fma(dz, dz, d2[b]); }
                    else if constexpr (DIM3) { const T dz = az - cj.z; d2[b] = t_fma(dz, dz, d2[b]); }
                }
                if constexpr (SWEEP) {
                    // distances only: they serve every parameter vector of the chunk
                    sqrt_batch<CB>(d2);
#pragma unroll
                    for (int b = 0; b < CB; ++b)
                        if (t0 + b < NP) Dst[t0 + b] = d2[b];
                    continue;
                }
                cov_batch<KERN, CB, TB, REP>(d2, exp_lane, sigma2);
#pragma unroll
                for (int b = 0; b < CB; ++b) {
                    if (t0 + b < NP) {
                        const SlotInfo si = slot_info<G, R, FOLD>(t0 + b);
                        A[si.s][si.j] = d2[b];
                        if (si.diag) A[si.s1][si.j1] = d2[b];  // a don't-care entry for the lanes that did not mean it
                    }
                }
            }
            set_diag(A, dg);
        }
        __syncwarp();  // tile and stage[] are rewritten by the next iteration

        if (EMIT && live) {
            // per-location covariance blocks (the _CNs/_Ccross/_Cs accessors; parity output)
            const int64_t o = i - a.lo;
#pragma unroll
            for (int s = 0; s < R; ++s) {
                const int r = rowq(s);
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    if (j >= (s + 1) * G || j > r) continue;
                    // a padded row or column reports exactly 0 (the build itself leaves ~1e-304 there: its exponent
                    // arithmetic caps u at 700); a padded column is recognised by its staged sentinel coordinate
                    const bool col_valid = j == r || double(stage[j].x) < (sizeof(T) == 8 ? 1e99 : 1e14);
                    const double v = valid[s] && col_valid ? double(A[s][j]) : 0.0;
                    if (r == P - 1) {
                        if (j < m && a.cc) a.cc[o * m + j] = v;
                        if (j == P - 1 && a.cs) a.cs[o] = v;
                    } else if (r < m && a.CN) {
                        a.CN[(o * m + r) * m + j] = v;
                        a.CN[(o * m + j) * m + r] = v;
                    }
                }
            }
        }

        // sweep: one pass of [covariances from the stored distances -> elimination -> accumulation] per
        // parameter vector of the chunk; otherwise a single pass over the matrix built above
        int kk = 0;
#pragma unroll 1
        do {
        volatile double *ab = accbuf + (SWEEP ? kk * 128 : 0);
        if constexpr (SWEEP) {
            constexpr int CBs = 6;
            constexpr int NP = n_slots<G, R, FOLD>();
            const T phik = T(sw_prm[kk][0]), dlt = T(sw_prm[kk][1]), is2 = T(sw_prm[kk][2]);
#pragma unroll
            for (int s = 0; s < R; ++s) {
                w[s] = w0[s];
                dg[s] = valid[s] ? t_fma(e2r[s], is2, T(1) + dlt) : T(1);  // 1 + (tau2 + eps2) / sigma2
            }
#pragma unroll
            for (int t0 = 0; t0 < NP; t0 += CBs) {
                T x[CBs];
#pragma unroll
                for (int b = 0; b < CBs; ++b) x[b] = Dst[(t0 + b < NP) ? t0 + b : NP - 1] * phik;
                corr_batch<KERN, CBs, TB, REP>(x, exp_lane);
#pragma unroll
                for (int b = 0; b < CBs; ++b) {
                    if (t0 + b < NP) {
                        const SlotInfo si = slot_info<G, R, FOLD>(t0 + b);
                        A[si.s][si.j] = x[b];
                        if (si.diag) A[si.s1][si.j1] = x[b];
                    }
                }
            }
            set_diag(A, dg);
        }

        // ---- stage 3: LDL^T elimination with the right-hand side carried along -----------------
        bool bad = false;
        T Flast = T(1), rlast = T(0);
        if constexpr (ELIM == 1) {
            using T2 = typename Pair2<T>::type;
#pragma unroll
            for (int k = 0; k < P; ++k) {
                T *col = (k & 1) ? col_odd : col_even;
                // publish column k (this lane's rows from the pivot's row block on) and w_k; two
                // buffers alternate, so one __syncwarp per pivot orders writes against earlier reads
#pragma unroll
                for (int s = 0; s < R; ++s)
                    if (s * G + G - 1 >= k) col[rowq(s)] = A[s][k];
                if (rowq(k / G) == k) col[P] = w[k / G];
                __syncwarp();
                // even k: the 16-byte load of the pair (k, k + 1) that the trailing update needs anyway carries D_k in its
                // first half (an 8-byte load of col[k] alone would be one more wavefront on the data stage)
                T2 first = {T(0), T(0)};
                if ((k & 1) == 0) first = *reinterpret_cast<const T2 *>(col + k);
                const T Dk = (k & 1) == 0 ? first.x : col[k], wk = col[P];
                bad |= !(Dk > T(0));
                if (k == P - 1) {
                    Flast = Dk;
                    rlast = wk;
                } else {
                    const T inv = fast_rcp(Dk);
                    T l[R];
#pragma unroll
                    for (int s = 0; s < R; ++s) {
                        if (s * G + G - 1 > k) {
                            l[s] = A[s][k] * inv;
                            w[s] = t_fma(-l[s], wk, w[s]);
                        }
                    }
#pragma unroll
                    for (int j0 = 0; j0 < P; j0 += 2) {
                        if (j0 + 1 > k) {
                            const T2 uu = j0 == k ? first : *reinterpret_cast<const T2 *>(col + j0);  // broadcast inside the group
                            if (j0 > k) {
#pragma unroll
                                for (int s = 0; s < R; ++s)
                                    if (s * G + G - 1 >= j0) A[s][j0] = t_fma(-l[s], uu.x, A[s][j0]);
                            }
#pragma unroll
                            for (int s = 0; s < R; ++s)
                                if (s * G + G - 1 >= j0 + 1) A[s][j0 + 1] = t_fma(-l[s], uu.y, A[s][j0 + 1]);
                        }
                    }
#pragma unroll
                    for (int s = 0; s < R; ++s)
                        if (s * G + G - 1 > k) A[s][k] = l[s];  // unit-lower factor (emit path)
                }
            }
        } else {
#pragma unroll
        for (int k = 0; k < P; ++k) {
            const T Dk = grp_bcast<T, G>(A[k / G][k], k % G);
            const T wk = grp_bcast<T, G>(w[k / G], k % G);
            bad |= !(Dk > T(0));
            if (k == P - 1) {
                Flast = Dk;
                rlast = wk;
            } else {
                const T inv = fast_rcp(Dk);
                T l[R];
#pragma unroll
                for (int s = 0; s < R; ++s) {
                    if (s * G + G - 1 > k) {
                        l[s] = A[s][k] * inv;
                        w[s] = t_fma(-l[s], wk, w[s]);
                    }
                }
                // (constant trip counts + guards: nvcc unrolls inner loops before the outer index is
                // known, and a k-dependent bound would leave a rolled remainder that indexes A[]
                // dynamically and demotes it to local memory)
#pragma unroll
                for (int j = 1; j < P; ++j) {
                    if (j > k) {
                        const T uj = grp_bcast<T, G>(A[j / G][k], j % G);
#pragma unroll
                        for (int s = 0; s < R; ++s)
                            if (s * G + G - 1 >= j) A[s][j] = t_fma(-l[s], uj, A[s][j]);
                    }
                }
#pragma unroll
                for (int s = 0; s < R; ++s)
                    if (s * G + G - 1 > k) A[s][k] = l[s];  // unit-lower factor (emit path)
            }
        }

        }

        if (live && q == 0) {
            const double Fd = double(Flast);
            bad |= !(Fd < INFINITY);
            if (bad) {
                if constexpr (SWEEP) ab[96] += 1.0; else ++nbad;
            } else {
                const int hi = __double2hiint(Fd);
                if (hi >= 0x00100000) {
                    // F = mant * 2^e, mant in [1, 2); the running product stays in [1, 2) as well
                    const double mant = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(Fd));
                    const double pr = ab[0] * mant;
                    const int ph = __double2hiint(pr);
                    const int carry = (ph >> 20) - 1023;  // 0 or 1
                    ab[0] = __hiloint2double(ph - (carry << 20), __double2loint(pr));
                    ab[64] += double((hi >> 20) - 1023 + carry);
                } else {
                    ab[64] += log2(Fd);  // subnormal F: the exponent field is not usable
                }
                ab[32] = fma(double(rlast) * double(rlast), fast_rcp(Fd), ab[32]);
            }
        }

        if constexpr (EMIT) {
            // b_i = L_N^{-T} ell, ell = last row of the unit-lower factor.  One location at a time:
            // its G lanes dump their rows to shared memory and lane 0 back-substitutes (parity /
            // prediction output, not the metric's path).
            for (int gg = 0; gg < W; ++gg) {
                if (g == gg) {
#pragma unroll
                    for (int s = 0; s < R; ++s) {
                        const int r = rowq(s);
#pragma unroll
                        for (int kk = 0; kk < P; ++kk)
                            if (kk < (s + 1) * G && kk < r) dump[r * P + kk] = A[s][kk];
                    }
                }
                __syncwarp();
                if (g == gg && live && q == 0) {
                    const int64_t o = i - a.lo;
                    if (a.F) a.F[o] = bad ? nan("") : double(Flast);
                    if (a.B) {
                        double b[P];
                        for (int kk = P - 2; kk >= 0; --kk) {
                            double v = double(dump[(P - 1) * P + kk]);
                            for (int r = kk + 1; r < P - 1; ++r) v -= double(dump[r * P + kk]) * b[r];
                            b[kk] = v;
                        }
                        for (int kk = 0; kk < m; ++kk) a.B[o * m + kk] = bad ? nan("") : b[kk];
                    }
                }
                __syncwarp();
            }
        }
        if constexpr (SWEEP) __syncwarp();  // the column buffers are rewritten by the next parameter vector
        } while (SWEEP && ++kk < kc);  // parameter vectors (a single pass outside the sweep variant)
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");  // drain copies issued for groups past the end
    TL(3, blockIdx.x == 0 && threadIdx.x == 0);
#ifdef NNGP_TIMELINE
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); nngp_tl_blk[1][blockIdx.x] = g_; }
#endif

    // ---- reduction: warp shuffle tree -> block -> per-block partial -> last block sums ------
    // (one pass per parameter vector of the chunk; a single pass outside the sweep variant)
    __shared__ double red[kWarps][3];
    __shared__ bool is_last;
    for (int kk = 0; kk < kc; ++kk) {
        volatile double *ab = accbuf + (SWEEP ? kk * 128 : 0);
        double acc_log = fma(ab[64], 0.6931471805599453094, log(ab[0])), acc_quad = ab[32];
        double bad_d = SWEEP ? ab[96] : double(nbad);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            acc_log += __shfl_xor_sync(0xffffffffu, acc_log, off);
            acc_quad += __shfl_xor_sync(0xffffffffu, acc_quad, off);
            bad_d += __shfl_xor_sync(0xffffffffu, bad_d, off);
        }
        __syncthreads();  // red[] of the previous pass has been read
        if (lane == 0) { red[warp][0] = acc_log; red[warp][1] = acc_quad; red[warp][2] = bad_d; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double *part = a.partials + (size_t(k0 + kk) * gridDim.x + blockIdx.x) * 3;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            for (int wv = 0; wv < kWarps; ++wv) { s0 += red[wv][0]; s1 += red[wv][1]; s2 += red[wv][2]; }
            part[0] = s0; part[1] = s1; part[2] = s2;
        }
    }
    TL(4, blockIdx.x == 0 && threadIdx.x == 0);
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int ticket = atomicAdd(a.counters + blockIdx.y, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    TL(5, blockIdx.x == 0 && threadIdx.x == 0);
    if (is_last) {
        TL(6, threadIdx.x == 0);
        __threadfence();
        double(*fin)[3] = reinterpret_cast<double(*)[3]>(smem_raw);  // main-loop buffers are dead here
        for (int kk = 0; kk < kc; ++kk) {
            const double *part = a.partials + size_t(k0 + kk) * gridDim.x * 3;
            double t0 = 0.0, t1 = 0.0, t2 = 0.0;
            for (unsigned int b = threadIdx.x; b < gridDim.x; b += kThreads) {
                t0 += __ldcg(part + size_t(b) * 3 + 0);
                t1 += __ldcg(part + size_t(b) * 3 + 1);
                t2 += __ldcg(part + size_t(b) * 3 + 2);
            }
            double tot[3];
            // Both trees are fixed: deterministic for a given grid.  (Two forms on purpose: ptxas allocates the
            // whole kernel at once, and giving the sweep variant the shuffle tree below costs its main loop
            // 40 bytes of spill traffic per iteration -- 10 % of its run time.)
            if constexpr (SWEEP) {
                __syncthreads();
                fin[threadIdx.x][0] = t0; fin[threadIdx.x][1] = t1; fin[threadIdx.x][2] = t2;
                __syncthreads();
                for (int stride = kThreads / 2; stride > 0; stride >>= 1) {
                    if (threadIdx.x < stride) {
                        fin[threadIdx.x][0] += fin[threadIdx.x + stride][0];
                        fin[threadIdx.x][1] += fin[threadIdx.x + stride][1];
                        fin[threadIdx.x][2] += fin[threadIdx.x + stride][2];
                    }
                    __syncthreads();
                }
                tot[0] = fin[0][0]; tot[1] = fin[0][1]; tot[2] = fin[0][2];
            } else {  // one launch = one evaluation: the tail is part of the fixed cost, so it is kept short
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    t0 += __shfl_xor_sync(0xffffffffu, t0, off);
                    t1 += __shfl_xor_sync(0xffffffffu, t1, off);
                    t2 += __shfl_xor_sync(0xffffffffu, t2, off);
                }
                __syncthreads();  // the main loop's buffers are no longer read
                if ((threadIdx.x & 31) == 0) { fin[threadIdx.x >> 5][0] = t0; fin[threadIdx.x >> 5][1] = t1; fin[threadIdx.x >> 5][2] = t2; }
                __syncthreads();
                tot[0] = 0.0; tot[1] = 0.0; tot[2] = 0.0;
#pragma unroll
                for (int wv = 0; wv < kWarps; ++wv) { tot[0] += fin[wv][0]; tot[1] += fin[wv][1]; tot[2] += fin[wv][2]; }
            }
            if constexpr (FACT) {
                // restore sigma2: F = sigma2 F' on the (n - n_bad) locations that entered the sums
                const double good = double(a.hi - a.lo) - tot[2];
                double ls2, is2;
                if constexpr (SWEEP) {
                    ls2 = sw_prm[kk][3];
                    is2 = sw_prm[kk][2];
                } else {
                    // the parameter pointer is rebuilt from %ctaid here on purpose: carried from the top of the
                    // kernel it stays live across the main loop, which is at its register limit and spills for it
                    unsigned int by_;
                    asm volatile("mov.u32 %0, %%ctaid.y;" : "=r"(by_));
                    const double s2 = prm_at(int(by_), 0);
                    ls2 = log(s2);
                    is2 = 1.0 / s2;
                }
                tot[0] = fma(good, ls2, tot[0]);
                tot[1] *= is2;
                // sigma2 must be positive and finite to be factored out: otherwise every location counts as
                // bad (the contract is a count, never a NaN)
                if (!(is2 > 0.0 && is2 < INFINITY)) { tot[0] = 0.0; tot[1] = 0.0; tot[2] = double(a.hi - a.lo); }
            }
            if (a.px.world > 1) {  // sum over the ranks through NVLink peer memory (block-uniform branch)
                __syncthreads();
                peer_allreduce3(a.px, k0 + kk, tot[0], tot[1], tot[2], &fin[kThreads / 2][0], tot);
            }
            if (threadIdx.x == 0) publish_result(a, k0 + kk, tot);
        }
        if (threadIdx.x == 0) a.counters[blockIdx.y] = 0u;  // ready for the next launch
        TL(7, threadIdx.x == 0);
    }
}

// ---- host-side dispatch of one (T, KERN) family -------------------------------------------------
template <typename T, int G, int R, int KERN, bool DIM3, int MINB, int BUILD, int ELIM, int FOLD>
cudaError_t launch_one(const EvalArgs &a, int K, int grid_x, cudaStream_t stream)
{
    auto kern = fused_loglik_kernel<T, G, R, KERN, DIM3, MINB, BUILD, ELIM, false, FOLD>;
    if constexpr (!sweep_build<BUILD>())
        if (a.emit) kern = fused_loglik_kernel<T, G, R, KERN, DIM3, MINB, BUILD, ELIM, true, FOLD>;
    const size_t smem = smem_bytes<T, G, R, DIM3, BUILD, ELIM>(a.emit != 0);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
#ifdef NNGP_TUNE
    if (const char *c = getenv("NNGP_TUNE_CARVEOUT")) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(c));
#endif
    const int grid_y = sweep_build<BUILD>() ? (K + kSweepChunk - 1) / kSweepChunk : K;
    kern<<<dim3(grid_x, grid_y, 1), kThreads, smem, stream>>>(a);
    return cudaGetLastError();
}

template <typename T, int G, int R, int KERN, bool DIM3, int MINB, int BUILD, int ELIM, int FOLD>
int blocks_per_sm()
{
    auto kern = fused_loglik_kernel<T, G, R, KERN, DIM3, MINB, BUILD, ELIM, false, FOLD>;
    int nb = 0;
    const size_t smem = smem_bytes<T, G, R, DIM3, BUILD, ELIM>(false);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, smem) != cudaSuccess) nb = 1;
    return nb < 1 ? 1 : nb;
}

// ---- shape table: (G lanes per location, R rows per lane) -> P = G*R >= m + 1 rows --------------------
// fp64 (the folded layout, every lane evaluating exactly P(P-1)/2G pairs):
//   m <=  7 : (4, 2) folded, 4-pair batches
//   m <= 15 : (4, 4) folded, 6-pair batches, 255 registers, 8 warps/SM (in-thread parallelism hides the FP64 latency
//             better than a third block of warps: 0.432 ms at cfg3 against 0.514; unfolded 0.440; (8, 2) folded at
//             8 / 12 / 16 warps per SM 0.577 / 0.588 / 0.614 -- half the rows per lane means half the locations per
//             warp, and the per-iteration bookkeeping does not shrink with them)
//   m <= 31 : (16, 2) folded, 6-pair batches: 5.37 ms per 2e6 locations at m = 30 in 3-D against 6.02 for the unfolded
//             (8, 4) it replaces (folded (8, 4): 5.74); 2-D 4.78 against 5.14 for the rolled pair-list build
//   m == 32 : (16, 3) rolled pair-list build (three row blocks do not pair)
// fp32: (4, 2) / (4, 4) row owner, (8, 4) / (16, 3) rolled pair-list build, 16 warps/SM.
// MINB = resident blocks per SM the kernel is compiled for (the register cap).
struct ShapeInfo {
    int blocks_per_sm;
    int loc_per_warp;
};

constexpr int kElim = 1;  // production elimination variant (see fused_loglik_kernel)

template <typename T, int KERN>
struct Launcher {
    const EvalArgs &a;
    int K, grid_x;
    cudaStream_t stream;
    bool sweep() const { return K >= 2 && !a.emit; }
    template <int G, int R, bool DIM3, int MINB, int BUILD, int ELIM = kElim, int FOLD = 0>
    cudaError_t run() const { return launch_one<T, G, R, KERN, DIM3, MINB, BUILD, ELIM, FOLD>(a, K, grid_x, stream); }
};
template <typename T, int KERN>
struct Describer {
    bool sweep() const { return false; }  // grid sizing follows the single-vector kernel (same shape, same registers)
    template <int G, int R, bool DIM3, int MINB, int BUILD, int ELIM = kElim, int FOLD = 0>
    ShapeInfo run() const { return ShapeInfo{blocks_per_sm<T, G, R, KERN, DIM3, MINB, BUILD, ELIM, FOLD>(), 32 / G}; }
};

template <typename T, bool DIM3, typename F>
auto dispatch_shape(int m, const F &f)
{
    constexpr bool F64 = sizeof(T) == 8;
#ifdef NNGP_TUNE  // development knobs (results: DESIGN.md 5.3); never compiled into the shipped library
    if constexpr (F64) {
        if (const char *e = getenv("NNGP_TUNE_SHAPE"); e && m > 7 && m <= 15 && !f.sweep()) {
            if (!strcmp(e, "44")) return f.template run<4, 4, DIM3, 2, 2, 1, 0>();    // unfolded row owner (round 1)
            if (!strcmp(e, "44f")) return f.template run<4, 4, DIM3, 2, 2, 1, 1>();   // folded, 6-pair batches
            if (!strcmp(e, "44g")) return f.template run<4, 4, DIM3, 2, 2, 1, 2>();   // folded, row-block-major slots
            if (!strcmp(e, "44f4")) return f.template run<4, 4, DIM3, 2, 0, 1, 1>();  // folded, 4-pair batches
            if (!strcmp(e, "44f9")) return f.template run<4, 4, DIM3, 2, 3, 1, 1>();  // folded, 9-pair batches
            if (!strcmp(e, "44f3")) return f.template run<4, 4, DIM3, 3, 2, 1, 1>();  // folded, 6-pair batches, 12 warps/SM
            if (!strcmp(e, "44f34")) return f.template run<4, 4, DIM3, 3, 0, 1, 1>(); // folded, 4-pair batches, 12 warps/SM
            if (!strcmp(e, "82f3")) return f.template run<8, 2, DIM3, 3, 2, 1, 1>();  // 8 lanes x 2 rows, 12 warps/SM
            if (!strcmp(e, "28f")) return f.template run<2, 8, DIM3, 2, 2, 1, 1>();   // 2 lanes x 8 rows, 6-pair batches
            if (!strcmp(e, "28f4")) return f.template run<2, 8, DIM3, 2, 0, 1, 1>();  // 2 lanes x 8 rows, 4-pair batches
            if (!strcmp(e, "28g4")) return f.template run<2, 8, DIM3, 2, 0, 1, 2>();  // row-block-major slot order
        }
        if (const char *e = getenv("NNGP_TUNE_SHAPE"); e && m > 15 && m <= 31 && !f.sweep()) {
            if (!strcmp(e, "84")) return f.template run<8, 4, DIM3, 2, 0, 1, 0>();
            if (!strcmp(e, "84f")) return f.template run<8, 4, DIM3, 2, 0, 1, 1>();
            if (!strcmp(e, "84g")) return f.template run<8, 4, DIM3, 2, 0, 1, 2>();
            if (!strcmp(e, "84f6")) return f.template run<8, 4, DIM3, 2, 2, 1, 1>();  // 6-pair batches
            if (!strcmp(e, "84r")) return f.template run<8, 4, DIM3, 2, 1, 1, 0>();   // rolled pair-list build
            if (!strcmp(e, "162f")) return f.template run<16, 2, DIM3, 2, 2, 1, 1>();
            if (!strcmp(e, "162f3")) return f.template run<16, 2, DIM3, 3, 2, 1, 1>();
            if (!strcmp(e, "162f34")) return f.template run<16, 2, DIM3, 3, 0, 1, 1>();
            if (!strcmp(e, "162g")) return f.template run<16, 2, DIM3, 2, 2, 1, 2>();
            if (!strcmp(e, "162f4")) return f.template run<16, 2, DIM3, 2, 0, 1, 1>();
            if (!strcmp(e, "162g4")) return f.template run<16, 2, DIM3, 2, 0, 1, 2>();
            if (!strcmp(e, "162g3")) return f.template run<16, 2, DIM3, 3, 0, 1, 2>();
        }
    }
#endif
    if constexpr (F64) {  // K >= 2: the sweep variant (distances shared by the parameter vectors)
        if (m <= 7 && f.sweep()) return f.template run<4, 2, DIM3, 4, 6, kElim, 1>();
        if (m <= 15 && f.sweep()) return f.template run<4, 4, DIM3, 2, 6, kElim, 1>();
        if (m <= 31 && f.sweep()) return f.template run<16, 2, DIM3, 2, 6, kElim, 1>();

        // fp64 single vector: the folded layouts (measured, tools/tune.py: DESIGN.md 5.3)
        if (m <= 7) return f.template run<4, 2, DIM3, 4, 0, kElim, 1>();
        if (m <= 15) return f.template run<4, 4, DIM3, 2, 2, kElim, 1>();
        if (m <= 31) return f.template run<16, 2, DIM3, 2, 2, kElim, 1>();
        return f.template run<16, 3, DIM3, 2, 1>();
    } else {
        if (m <= 7) return f.template run<4, 2, DIM3, 4, 0, kElim, 1>();
        if (m <= 15) return f.template run<4, 4, DIM3, 4, 0, kElim, 1>();
        if (m <= 31) return f.template run<8, 4, DIM3, 4, 1>();
        return f.template run<16, 3, DIM3, 4, 1>();
    }
}

template <typename T, int KERN>
cudaError_t launch_family(int m, int D, const EvalArgs &a, int K, int grid_x, cudaStream_t stream)
{
    const Launcher<T, KERN> f{a, K, grid_x, stream};
    return D == 3 ? dispatch_shape<T, true>(m, f) : dispatch_shape<T, false>(m, f);
}

template <typename T, int KERN>
ShapeInfo shape_family(int m, int D)
{
    const Describer<T, KERN> f{};
    return D == 3 ? dispatch_shape<T, true>(m, f) : dispatch_shape<T, false>(m, f);
}

}  // namespace nngp_fused

// one translation unit per (dtype, kernel family) keeps nvcc parallel; each defines these two.
#define NNGP_DEFINE_FAMILY(NAME, T, KERN)                                                         \
    cudaError_t nngp_launch_##NAME(int m, int D, const EvalArgs &a, int K, int grid_x,            \
                                   cudaStream_t stream)                                           \
    {                                                                                             \
        return nngp_fused::launch_family<T, KERN>(m, D, a, K, grid_x, stream);                    \
    }                                                                                             \
    void nngp_shape_##NAME(int m, int D, int *blocks_per_sm, int *loc_per_warp)                   \
    {                                                                                             \
        const nngp_fused::ShapeInfo si = nngp_fused::shape_family<T, KERN>(m, D);                 \
        *blocks_per_sm = si.blocks_per_sm;                                                        \
        *loc_per_warp = si.loc_per_warp;                                                          \
    }
