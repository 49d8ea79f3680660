/*
 * nngp_oracle.c -- CPU ORACLE for the NNGP likelihood hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file.  The product path (pynngp_b200/) never links, imports or calls it.
 *
 * What it restates (reference = bwpriest/pyNNGP, citations are pyNNGP/nngp.py:LINE):
 *   stage 1  oracle_knn_ordered   <- _make_s_neighbor_sets, nngp.py:49-62.  The reference builds
 *            KDTree(s[0:i]) per i and queries k=min(m,i) (nngp.py:55-61); the arithmetic lives in
 *            scikit-learn (unpinned in setup.py:70; 1.9.0 installed here): squared Euclidean
 *            distance accumulated dimension by dimension in fp64 without FMA
 *            (sklearn/metrics/_dist_metrics.pxd.tp:39-49), neighbours returned in ascending
 *            distance.  Restated as an exact brute-force scan with the total order (d2, j).
 *            PINNED: tests/golden/ns_*.npz were produced by running the unmodified reference
 *            (tests/golden/make_golden.py) and this function reproduces them bit for bit.
 *   stage 2  oracle_CNs / oracle_Ccross / oracle_Cs  <- _CNs nngp.py:78-82, _Ccross nngp.py:84-86,
 *            _Cs nngp.py:92-96 (docstrings C_{N(s_i)}, C_{s_i,N(s_i)}, C_{si,si}).
 *   stage 3  oracle_Bsi / oracle_Fsi <- _Bsi nngp.py:73-76 (B_{s_i}), _Fsi nngp.py:88-90 (F_{s_i});
 *            oracle_loglik <- the reduction BASELINE.json north_star defines:
 *            sum_i [ log F_i + (y_i - b_i^T y_N(i))^2 / F_i ].
 *            PARITY UNPINNED BY THE REFERENCE: nngp.py:73-96 are empty stubs and the reference has no
 *            test vector for them.  The oracle is anchored instead to closed-form known answers
 *            (dense-GP identity: with m >= n-1 the NNGP density equals the exact multivariate normal
 *            density; i=0 and p=1 hand cases) in tests/test_oracle.py.
 *
 * Covariance model (the reference leaves `cov` to the caller, nngp.py:12; this is the
 * parametrisation the engine and the oracle share):
 *   C(a,b) = sigma2 * rho(phi * ||s_a - s_b||)            a != b
 *   C(a,a) = sigma2 + tau2 + eps2[a]
 *   kernel 0 exponential   rho(u) = exp(-u)
 *   kernel 1 Matern nu=3/2 rho(u) = (1 + u) exp(-u)
 *   kernel 2 Matern nu=5/2 rho(u) = (1 + u + u^2/3) exp(-u)
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off).  Single-threaded C; callers that want all
 * host cores split [lo, hi) into chunks and call from threads (ctypes releases the GIL) -- this
 * image has no libgomp.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAX_M 64

typedef struct {
    const double *coords; /* n x D row-major */
    const double *y;      /* n */
    const double *eps2;   /* n or NULL */
    const int32_t *nbr;   /* n x m row-major, -1 padded */
    int64_t n;
    int D;
    int m;
    int kernel_id;
    double sigma2, phi, tau2;
} oracle_ctx;

/* squared distance exactly as sklearn's euclidean_rdist: d = 0; d += t*t per dimension, no FMA */
static double dist2(const double *a, const double *b, int D)
{
    double d = 0.0;
    for (int k = 0; k < D; ++k) {
        double t = a[k] - b[k];
        d += t * t; /* built with -ffp-contract=off: no FMA contraction */
    }
    return d;
}

/* ---- stage 1: ordered k-NN, nngp.py:49-62 ------------------------------------------------ */
/* out: n x m int32, row i holds the min(m,i) nearest predecessors j<i in ascending (d2, j),
 * padded with -1.  rows [lo, hi) only (others untouched). */
void oracle_knn_ordered(const double *coords, int64_t n, int D, int m, int64_t lo, int64_t hi,
                        int32_t *out)
{
    for (int64_t i = lo; i < hi; ++i) {
        double bd[ORACLE_MAX_M];
        int32_t bj[ORACLE_MAX_M];
        int cnt = 0;
        const double *si = coords + i * D;
        for (int64_t j = 0; j < i; ++j) {
            double d2 = dist2(si, coords + j * D, D);
            if (cnt == m && !(d2 < bd[m - 1])) continue; /* ties at the boundary: smaller j stays */
            int pos = cnt < m ? cnt : m - 1;
            while (pos > 0 && d2 < bd[pos - 1]) { /* strict: equal d2 keeps the earlier j first */
                bd[pos] = bd[pos - 1];
                bj[pos] = bj[pos - 1];
                --pos;
            }
            bd[pos] = d2;
            bj[pos] = (int32_t)j;
            if (cnt < m) ++cnt;
        }
        int32_t *row = out + i * (int64_t)m;
        for (int k = 0; k < m; ++k) row[k] = k < cnt ? bj[k] : -1;
    }
}

/* ---- covariance ---------------------------------------------------------------------------- */
static double corr(int kernel_id, double u)
{
    switch (kernel_id) {
    case 0: return exp(-u);
    case 1: return (1.0 + u) * exp(-u);
    default: return (1.0 + u + u * u / 3.0) * exp(-u);
    }
}

static double cov_offdiag(const oracle_ctx *c, int64_t a, int64_t b)
{
    double d = sqrt(dist2(c->coords + a * c->D, c->coords + b * c->D, c->D));
    return c->sigma2 * corr(c->kernel_id, c->phi * d);
}

/* _Cs(i), nngp.py:92-96: C(s_i, s_i) */
double oracle_Cs(const oracle_ctx *c, int64_t i)
{
    return c->sigma2 + c->tau2 + (c->eps2 ? c->eps2[i] : 0.0);
}

static int nbr_count(const oracle_ctx *c, int64_t i)
{
    int p = 0;
    const int32_t *row = c->nbr + i * (int64_t)c->m;
    while (p < c->m && row[p] >= 0) ++p;
    return p;
}

/* _CNs(i), nngp.py:78-82: C_{N(s_i)}, p x p row-major into CN (leading dimension m). returns p */
int oracle_CNs(const oracle_ctx *c, int64_t i, double *CN)
{
    int p = nbr_count(c, i);
    const int32_t *row = c->nbr + i * (int64_t)c->m;
    for (int a = 0; a < p; ++a)
        for (int b = 0; b < p; ++b)
            CN[a * c->m + b] = (a == b) ? oracle_Cs(c, row[a]) : cov_offdiag(c, row[a], row[b]);
    return p;
}

/* _Ccross(i), nngp.py:84-86: C_{s_i, N(s_i)}, length p. returns p */
int oracle_Ccross(const oracle_ctx *c, int64_t i, double *cc)
{
    int p = nbr_count(c, i);
    const int32_t *row = c->nbr + i * (int64_t)c->m;
    for (int a = 0; a < p; ++a) cc[a] = cov_offdiag(c, i, row[a]);
    return p;
}

/* in-place lower Cholesky of the leading p x p block (ld = m). returns 0 ok, 1 not SPD */
static int chol(double *A, int p, int ld)
{
    for (int k = 0; k < p; ++k) {
        double s = A[k * ld + k];
        for (int t = 0; t < k; ++t) s -= A[k * ld + t] * A[k * ld + t];
        if (!(s > 0.0)) return 1;
        double l = sqrt(s);
        A[k * ld + k] = l;
        for (int r = k + 1; r < p; ++r) {
            double v = A[r * ld + k];
            for (int t = 0; t < k; ++t) v -= A[r * ld + t] * A[k * ld + t];
            A[r * ld + k] = v / l;
        }
    }
    return 0;
}

/* _Bsi(i), nngp.py:73-76: b_i = C_N(i)^{-1} c_i  (length p, zero-padded to m)
 * _Fsi(i), nngp.py:88-90: F_i = C(i,i) - c_i^T b_i
 * returns p, or -1 if C_N(i) is not positive definite */
int oracle_Bsi_Fsi(const oracle_ctx *c, int64_t i, double *b, double *F)
{
    double CN[ORACLE_MAX_M * ORACLE_MAX_M];
    double cc[ORACLE_MAX_M], z[ORACLE_MAX_M];
    int m = c->m;
    int p = oracle_CNs(c, i, CN);
    oracle_Ccross(c, i, cc);
    for (int a = 0; a < m; ++a) b[a] = 0.0;
    if (chol(CN, p, m)) { *F = NAN; return -1; }
    for (int a = 0; a < p; ++a) { /* L z = c */
        double v = cc[a];
        for (int t = 0; t < a; ++t) v -= CN[a * m + t] * z[t];
        z[a] = v / CN[a * m + a];
    }
    for (int a = p - 1; a >= 0; --a) { /* L^T b = z */
        double v = z[a];
        for (int t = a + 1; t < p; ++t) v -= CN[t * m + a] * b[t];
        b[a] = v / CN[a * m + a];
    }
    double dot = 0.0;
    for (int a = 0; a < p; ++a) dot += cc[a] * b[a];
    *F = oracle_Cs(c, i) - dot;
    return p;
}

/* ---- flat entry points for ctypes ----------------------------------------------------------- */
static void fill_ctx(oracle_ctx *c, const double *coords, const double *y, const double *eps2,
                     const int32_t *nbr, int64_t n, int D, int m, int kernel_id,
                     const double *params)
{
    c->coords = coords; c->y = y; c->eps2 = eps2; c->nbr = nbr;
    c->n = n; c->D = D; c->m = m; c->kernel_id = kernel_id;
    c->sigma2 = params[0]; c->phi = params[1]; c->tau2 = params[2];
}

/* north_star reduction over rows [lo, hi): out[0] = sum log F_i, out[1] = sum r_i^2 / F_i,
 * out[2] = number of locations whose C_N(i) was not SPD or F_i <= 0 (excluded from the sums). */
void oracle_loglik(const double *coords, const double *y, const double *eps2, const int32_t *nbr,
                   int64_t n, int D, int m, int kernel_id, const double *params, int64_t lo,
                   int64_t hi, double *out)
{
    oracle_ctx c;
    fill_ctx(&c, coords, y, eps2, nbr, n, D, m, kernel_id, params);
    double slog = 0.0, squad = 0.0, bad = 0.0;
    for (int64_t i = lo; i < hi; ++i) {
        double b[ORACLE_MAX_M], F;
        int p = oracle_Bsi_Fsi(&c, i, b, &F);
        if (p < 0 || !(F > 0.0)) { bad += 1.0; continue; }
        const int32_t *row = nbr + i * (int64_t)m;
        double r = y[i];
        for (int a = 0; a < p; ++a) r -= b[a] * y[row[a]];
        slog += log(F);
        squad += r * r / F;
    }
    out[0] = slog; out[1] = squad; out[2] = bad;
}

/* B: (hi-lo) x m, F: (hi-lo) */
void oracle_factors(const double *coords, const double *y, const double *eps2, const int32_t *nbr,
                    int64_t n, int D, int m, int kernel_id, const double *params, int64_t lo,
                    int64_t hi, double *B, double *F)
{
    oracle_ctx c;
    fill_ctx(&c, coords, y, eps2, nbr, n, D, m, kernel_id, params);
    for (int64_t i = lo; i < hi; ++i)
        oracle_Bsi_Fsi(&c, i, B + (i - lo) * (int64_t)m, F + (i - lo));
}

/* CN: (hi-lo) x m x m (zero outside the leading p x p), cc: (hi-lo) x m, cs: (hi-lo) */
void oracle_cov_blocks(const double *coords, const double *eps2, const int32_t *nbr, int64_t n,
                       int D, int m, int kernel_id, const double *params, int64_t lo, int64_t hi,
                       double *CN, double *cc, double *cs)
{
    oracle_ctx c;
    fill_ctx(&c, coords, NULL, eps2, nbr, n, D, m, kernel_id, params);
    for (int64_t i = lo; i < hi; ++i) {
        double *cn = CN + (i - lo) * (int64_t)m * m;
        double *cr = cc + (i - lo) * (int64_t)m;
        memset(cn, 0, sizeof(double) * m * m);
        memset(cr, 0, sizeof(double) * m);
        oracle_CNs(&c, i, cn);
        oracle_Ccross(&c, i, cr);
        cs[i - lo] = oracle_Cs(&c, i);
    }
}

int oracle_max_m(void) { return ORACLE_MAX_M; }
