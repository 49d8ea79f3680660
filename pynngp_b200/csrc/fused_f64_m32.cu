// fused_f64_m32.cu -- instantiates the fused covariance/factorisation/reduction kernel
// (loglik_fused.cuh) for arithmetic type double and correlation family NNGP_MATERN32.
#define NNGP_TUNE 1
#include "loglik_fused.cuh"

NNGP_DEFINE_FAMILY(f64_m32, double, NNGP_MATERN32)
