// fused_f32_m52.cu -- instantiates the fused covariance/factorisation/reduction kernel
// (loglik_fused.cuh) for arithmetic type float and correlation family NNGP_MATERN52.
#include "loglik_fused.cuh"

NNGP_DEFINE_FAMILY(f32_m52, float, NNGP_MATERN52)
