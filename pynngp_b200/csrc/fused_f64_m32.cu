// fused_f64_m32.cu -- instantiates the fused covariance/factorisation/reduction kernel
// (loglik_fused.cuh) for arithmetic type double and correlation family NNGP_MATERN32.
// Development builds only (never the shipped library): `NNGP_DEV_DEFINES="NNGP_TUNE" python -m pynngp_b200.build --force`
// compiles the shape knobs tools/tune.py drives through NNGP_TUNE_SHAPE; NNGP_TIMELINE adds per-phase / per-block
// globaltimer stamps (tools/timeline*.py).  Without those defines the dispatch reads no environment variable.
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
