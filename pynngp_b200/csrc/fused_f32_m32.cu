// fused_f32_m32.cu -- instantiates the fused covariance/factorisation/reduction kernel
// (loglik_fused.cuh) for arithmetic type float and correlation family NNGP_MATERN32.
#include "loglik_fused.cuh"

NNGP_DEFINE_FAMILY(f32_m32, float, NNGP_MATERN32)
