// Host-side entry points of the HBM-resident dense path (gpet_dense.cu) used by the posterior / final-fit wrappers when a
// training set is larger than the shared-memory kernels hold (mmax > GPET_MAX_TRAIN).
#pragma once
#include "gpet_common.cuh"

namespace gpet {

// bytes of `work` the large-training-set posterior needs (ncols = n for the full covariance, rp for the low-rank form)
int64_t posterior_big_workspace_bytes(int B, int mmax, int ncols);

// gpet_posterior_full_f64 for mmax > GPET_MAX_TRAIN: same arguments, same outputs
int posterior_big_full(const int32_t* xi, const double* y, const double* w, const int32_t* m, int mmax, int m_cap, int B, int n,
                       const double* sigma_f, double noise_y, double gp_alpha, const double* kd, double* mean, double* ys,
                       double* cov, int32_t* status, void* work, cudaStream_t st);

// gpet_posterior_lowrank_f64 for mmax > GPET_MAX_TRAIN
int posterior_big_lowrank(const int32_t* xi, const double* y, const double* w, const int32_t* m, int mmax, int m_cap, int B,
                          int n, const double* sigma_f, double noise_y, double gp_alpha, const double* kd, const double* Ur,
                          const double* lam, int rp, double* mean, double* ys, double* Mr, int32_t* status, void* work,
                          cudaStream_t st);

}  // namespace gpet
