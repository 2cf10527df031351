// Unit GP kernels of the final fit (sklearn kernels.py RBF / Matern __call__ and eval_gradient), shared by the
// shared-memory objective kernels (gpet_finalfit.cu) and the HBM-resident large-training-set path (gpet_dense.cu).
#pragma once
#include "gpet_common.cuh"

namespace gpet {

// kind: 0 RBF, 1 Matern nu=0.5, 2 Matern nu=1.5, 3 Matern nu=2.5.  D = squared scaled distance.
__device__ __forceinline__ double kern_val(int kind, double D) {
    if (kind == 0) return exp(-0.5 * D);
    const double d = sqrt(D);
    if (kind == 1) return exp(-d);
    if (kind == 2) { const double t = d * 1.7320508075688772; return (1.0 + t) * exp(-t); }
    const double t = d * 2.23606797749979;
    return (1.0 + t + t * t / 3.0) * exp(-t);
}
// d k / d log(length_scale) (sklearn kernels.py RBF/Matern eval_gradient)
__device__ __forceinline__ double kern_dlogl(int kind, double D) {
    if (kind == 0) return exp(-0.5 * D) * D;
    if (kind == 1) { const double d = sqrt(D); return d > 0.0 ? exp(-d) * d : 0.0; }
    if (kind == 2) return 3.0 * D * exp(-sqrt(3.0 * D));
    const double t = sqrt(5.0 * D);
    return 5.0 / 3.0 * D * (t + 1.0) * exp(-t);
}

// value and d/dlog(length_scale) with one exponential
__device__ __forceinline__ void kern_both(int kind, double D, double& k, double& dk) {
    if (kind == 0) { k = exp(-0.5 * D); dk = k * D; return; }
    const double d = sqrt(D);
    if (kind == 1) { k = exp(-d); dk = k * d; return; }
    if (kind == 2) { const double t = d * 1.7320508075688772; const double e = exp(-t); k = (1.0 + t) * e; dk = 3.0 * D * e; return; }
    const double t = d * 2.23606797749979;
    const double e = exp(-t);
    k = (1.0 + t + t * t / 3.0) * e;
    dk = 5.0 / 3.0 * D * (t + 1.0) * e;
}

}  // namespace gpet
