// fused_f64_exp.cu -- instantiates the fused covariance/factorisation/reduction kernel
// (loglik_fused.cuh) for arithmetic type double and correlation family NNGP_EXPONENTIAL.
#include "loglik_fused.cuh"

NNGP_DEFINE_FAMILY(f64_exp, double, NNGP_EXPONENTIAL)
