// fused_f32_exp.cu -- instantiates the fused covariance/factorisation/reduction kernel
// (loglik_fused.cuh) for arithmetic type float and correlation family NNGP_EXPONENTIAL.
#include "loglik_fused.cuh"

NNGP_DEFINE_FAMILY(f32_exp, float, NNGP_EXPONENTIAL)
