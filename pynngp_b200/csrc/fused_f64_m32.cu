// fused_f64_m32.cu -- instantiates the fused covariance/factorisation/reduction kernel
// (loglik_fused.cuh) for arithmetic type double and correlation family NNGP_MATERN32.
#define NNGP_TUNE 1
// #define NNGP_TIMELINE 1   // development: per-phase / per-block globaltimer stamps (tools/timeline*.py)
#include "loglik_fused.cuh"

NNGP_DEFINE_FAMILY(f64_m32, double, NNGP_MATERN32)

#ifdef NNGP_TIMELINE
extern "C" int nngp_debug_timeline(unsigned long long *out)
{
    return (int)cudaMemcpyFromSymbol(out, nngp_fused::nngp_tl, sizeof(unsigned long long) * 64);
}
extern "C" int nngp_debug_timeline_blocks(unsigned long long *out)
{
    return (int)cudaMemcpyFromSymbol(out, nngp_fused::nngp_tl_blk, sizeof(unsigned long long) * 3072);
}
#endif
