// fused_f64_m52.cu -- instantiates the fused covariance/factorisation/reduction kernel
// (loglik_fused.cuh) for arithmetic type double and correlation family NNGP_MATERN52.
#include "loglik_fused.cuh"

NNGP_DEFINE_FAMILY(f64_m52, double, NNGP_MATERN52)
