// Loop-carried state of GP_Edge_Tracing.__call__ (gpet.py:829-870) kept on the device: the observation sets, the
// decaying score threshold and the iteration counters of every trace, so that an iteration needs no host arithmetic
// and no per-trace device->host traffic (the host reads one 16-byte control block per iteration).
//
//   update_obs_kernel      compute_new_obs (gpet.py:589-616) on the per-bin maxima of gpet_select_f64: the threshold
//                          decay loop and the new observation set of every active trace
//   compact_active_kernel  which traces are still inside the while-loop (gpet.py:829), compacted to the front
//   training_sets_kernel   fit_predict_GP's training-set assembly (gpet.py:209-224) for the next iteration
#include "gpet_common.cuh"

namespace gpet {

constexpr int CT = 128;          // threads per CTA of the per-trace kernels
constexpr int MAX_DECAYS = 4000; // the reference loops forever when no threshold yields enough bins; we stop and flag

// ctrl[0] = number of active traces, ctrl[1] = error code (0 none, 1 Cholesky failed, 2 threshold loop cannot end),
// ctrl[2] = a trace that raised it, ctrl[3] = the largest observation count among the active traces
__device__ __forceinline__ void flag_error(int32_t* ctrl, int code, int trace) {
    if (atomicCAS(&ctrl[1], 0, code) == 0) ctrl[2] = trace;
}

// One CTA per active slot k (trace r = rows[k]).  smem: best[nb] f64 | old[max_old][2] i32 | keep[nb] i32
__global__ void __launch_bounds__(CT) update_obs_kernel(const double* __restrict__ bin_score,
                                                        const int32_t* __restrict__ bin_pos,
                                                        const int32_t* __restrict__ rows,
                                                        const int32_t* __restrict__ post_status, int nb, int N,
                                                        int max_old, int pixel_thresh, int algo_thresh,
                                                        int32_t* __restrict__ obs_xy, int32_t* __restrict__ n_obs,
                                                        double* __restrict__ thr, int32_t* __restrict__ n_iter,
                                                        int32_t* __restrict__ ctrl) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* best = reinterpret_cast<double*>(smem_raw);
    int32_t* old = reinterpret_cast<int32_t*>(best + nb);
    int32_t* keep = old + 2 * max_old;
    __shared__ double s_thr;
    __shared__ int s_warp[CT / 32];
    __shared__ int s_fail, s_base, s_have;
    const int k = blockIdx.x, tid = threadIdx.x;
    const int r = rows[k];
    if (post_status != nullptr && post_status[k] != 0) {
        if (tid == 0) flag_error(ctrl, 1, r);
        return;
    }
    const int n_pre = n_obs[r];
    for (int b = tid; b < nb; b += CT) best[b] = bin_score[(size_t)k * nb + b];
    for (int j = tid; j < 2 * n_pre; j += CT) old[j] = obs_xy[(size_t)r * max_old * 2 + j];
    if (tid == 0) { s_fail = 0; s_have = 0; }
    __syncthreads();
    // The loop of gpet.py:591-609 multiplies the threshold by 0.95 (by 1.0 on its first pass) until the number of bins
    // whose best score reaches it is >= T = min(n_pre + pixel_thresh, algo_thresh).  That count is monotone in the
    // threshold, so the loop ends at the first element of thr, thr*0.95, (thr*0.95)*0.95, ... that is <= v_T, the T-th
    // largest bin maximum (rank by counting; ties broken by bin index).
    const int T = min(n_pre + pixel_thresh, algo_thresh);
    for (int b = tid; b < nb; b += CT) {
        const double v = best[b];
        int rank = 0;
        for (int j = 0; j < nb; ++j) {
            const double u = best[j];
            rank += (u > v) || (u == v && j < b);
        }
        if (rank == T - 1) { s_thr = v; s_have = 1; }   // the ranks are a permutation: exactly one bin, when T <= nb
    }
    __syncthreads();
    if (tid == 0) {
        const double vt = s_thr;
        if (!s_have || !(vt > 0.0)) s_fail = 1;          // fewer non-empty bins than needed: the reference never ends
        else {
            double t = thr[r];
            int i = 0;
            while (t > vt && i < MAX_DECAYS) { t = __dmul_rn(t, 0.95); ++i; }   // gpet.py:595
            if (t > vt) s_fail = 1;
            s_thr = t;
        }
    }
    __syncthreads();
    if (s_fail) {
        if (tid == 0) flag_error(ctrl, 2, r);
        return;
    }
    const double t_fin = s_thr;
    // accepted bins in ascending order (np.unique, gpet.py:607, 613-616) -> new observation list
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += CT) {
        const int b = b0 + tid;
        const bool acc = b < nb && best[b] >= t_fin && best[b] >= 0.0;
        const unsigned bal = __ballot_sync(0xffffffffu, acc);
        if ((tid & 31) == 0) s_warp[tid >> 5] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < (tid >> 5); ++w) off += s_warp[w];
        off += __popc(bal & ((1u << (tid & 31)) - 1u));
        if (acc) keep[off] = b;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < CT / 32; ++w) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
    }
    const int n_new = s_base;
    for (int j = tid; j < n_new; j += CT) {
        const int p = bin_pos[(size_t)k * nb + keep[j]];
        int x, y;
        if (p < max_old) { x = old[2 * p]; y = old[2 * p + 1]; }       // an old observation keeps its bin
        else { const int q = p - max_old; x = q % N; y = q / N; }     // a new pixel (row-major index)
        obs_xy[((size_t)r * max_old + j) * 2] = x;
        obs_xy[((size_t)r * max_old + j) * 2 + 1] = y;
    }
    if (tid == 0) {
        n_obs[r] = n_new;
        thr[r] = t_fin;
        n_iter[r] += 1;
    }
}

// Single CTA: rows[0 .. n_active) = traces with n_obs < algo_thresh in ascending order; ctrl[0] = n_active.
__global__ void __launch_bounds__(1024) compact_active_kernel(const int32_t* __restrict__ n_obs, int B, int algo_thresh,
                                                              int32_t* __restrict__ rows, int32_t* __restrict__ ctrl,
                                                              int bump_iter) {
    __shared__ int s_warp[32];
    __shared__ int s_base, s_most;
    const int tid = threadIdx.x;
    if (tid == 0) s_base = s_most = 0;
    __syncthreads();
    for (int b0 = 0; b0 < B; b0 += 1024) {
        const int b = b0 + tid;
        const int no = b < B ? n_obs[b] : 0;
        const bool act = b < B && no < algo_thresh;
        if (act && no > 0) atomicMax(&s_most, no);
        const unsigned bal = __ballot_sync(0xffffffffu, act);
        if ((tid & 31) == 0) s_warp[tid >> 5] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < (tid >> 5); ++w) off += s_warp[w];
        off += __popc(bal & ((1u << (tid & 31)) - 1u));
        if (act) rows[off] = b;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < 32; ++w) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
    }
    if (tid == 0) {
        ctrl[0] = s_base;
        ctrl[3] = s_most;
    }
}

// One CTA per slot k < ctrl[0]: training set of trace rows[k] = stable sort by x of concat(init, obs) with the noise
// weights [alpha_init..., 1...] (gpet.py:209-214), plus the old observations in (row, col) order for gpet_select_f64.
__global__ void __launch_bounds__(CT) training_sets_kernel(const int32_t* __restrict__ init_xy,
                                                           const double* __restrict__ alpha_init, int K,
                                                           const int32_t* __restrict__ obs_xy,
                                                           const int32_t* __restrict__ n_obs, int max_old, int x_st,
                                                           int mmax, const int32_t* __restrict__ rows,
                                                           const int32_t* __restrict__ ctrl, int32_t* __restrict__ xi,
                                                           double* __restrict__ y, double* __restrict__ w,
                                                           int32_t* __restrict__ m_out, int32_t* __restrict__ old_yx,
                                                           int32_t* __restrict__ n_old) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t* px = reinterpret_cast<int32_t*>(smem_raw);   // [mmax]
    int32_t* py = px + mmax;                              // [mmax]
    const int k = blockIdx.x, tid = threadIdx.x;
    if (k >= ctrl[0]) return;
    const int r = rows[k];
    const int no = n_obs[r];
    const int m = K + no;
    for (int i = tid; i < m; i += CT) {
        const int32_t* src = i < K ? init_xy + ((size_t)r * K + i) * 2 : obs_xy + ((size_t)r * max_old + (i - K)) * 2;
        px[i] = src[0];
        py[i] = src[1];
    }
    __syncthreads();
    for (int i = tid; i < mmax; i += CT) {
        if (i < m) {
            const int xv = px[i];
            int rank = 0;
            for (int j = 0; j < m; ++j) rank += (px[j] < xv) || (px[j] == xv && j < i);
            xi[(size_t)k * mmax + rank] = xv - x_st;
            y[(size_t)k * mmax + rank] = (double)py[i];
            w[(size_t)k * mmax + rank] = i < K ? alpha_init[i] : 1.0;
        } else {
            xi[(size_t)k * mmax + i] = 0;
            y[(size_t)k * mmax + i] = 0.0;
            w[(size_t)k * mmax + i] = 0.0;
        }
    }
    for (int j = tid; j < max_old; j += CT) {
        const bool v = j < no;
        old_yx[((size_t)k * max_old + j) * 2] = v ? py[K + j] : 0;       // (row, col): gpet.py:857 swaps the columns
        old_yx[((size_t)k * max_old + j) * 2 + 1] = v ? px[K + j] : 0;
    }
    if (tid == 0) {
        m_out[k] = m;
        n_old[k] = no;
    }
}

}  // namespace gpet

using namespace gpet;

extern "C" int gpet_update_obs_f64(const double* bin_score, const int32_t* bin_pos, const int32_t* rows,
                                   const int32_t* post_status, int B_active, int nb, int N, int max_old,
                                   int pixel_thresh, int algo_thresh, int32_t* obs_xy, int32_t* n_obs, double* thr,
                                   int32_t* n_iter, int32_t* ctrl, void* stream) {
    GPET_REQUIRE(B_active >= 0 && nb > 0 && N > 0 && max_old >= nb, "gpet_update_obs_f64: bad sizes (max_old must be >= nb)");
    GPET_REQUIRE(pixel_thresh > 0 && algo_thresh > 0, "gpet_update_obs_f64: thresholds must be positive");
    if (B_active == 0) return GPET_OK;
    const size_t smem = (size_t)nb * 8 + (size_t)max_old * 8 + (size_t)nb * 4;
    GPET_SUPPORTED(smem <= 200 * 1024, "gpet_update_obs_f64: %d bins / %d observations exceed shared memory", nb, max_old);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(update_obs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    update_obs_kernel<<<B_active, CT, smem, (cudaStream_t)stream>>>(bin_score, bin_pos, rows, post_status, nb, N, max_old,
                                                                     pixel_thresh, algo_thresh, obs_xy, n_obs, thr,
                                                                     n_iter, ctrl);
    return check_launch("update_obs_kernel");
}

extern "C" int gpet_training_sets_f64(const int32_t* init_xy, const double* alpha_init, int K, const int32_t* obs_xy,
                                      const int32_t* n_obs, int B, int B_launch, int max_old, int algo_thresh, int x_st,
                                      int mmax, int bump_iter, int32_t* rows, int32_t* ctrl, int32_t* xi, double* y,
                                      double* w, int32_t* m, int32_t* old_yx, int32_t* n_old, int32_t* ctrl_host,
                                      void* stream) {
    GPET_REQUIRE(B > 0 && K > 0 && max_old >= 0 && mmax >= K + 1, "gpet_training_sets_f64: bad sizes");
    GPET_REQUIRE(B_launch >= 0 && B_launch <= B, "gpet_training_sets_f64: B_launch must be in 0..B");
    cudaStream_t st = (cudaStream_t)stream;
    compact_active_kernel<<<1, 1024, 0, st>>>(n_obs, B, algo_thresh, rows, ctrl, bump_iter);
    int rc = check_launch("compact_active_kernel");
    if (rc) return rc;
    if (B_launch > 0) {
        const size_t smem = (size_t)mmax * 8;
        GPET_SUPPORTED(smem <= 200 * 1024, "gpet_training_sets_f64: %d training points exceed shared memory", mmax);
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(training_sets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        training_sets_kernel<<<B_launch, CT, smem, st>>>(init_xy, alpha_init, K, obs_xy, n_obs, max_old, x_st, mmax, rows,
                                                         ctrl, xi, y, w, m, old_yx, n_old);
        rc = check_launch("training_sets_kernel");
        if (rc) return rc;
    }
    if (ctrl_host != nullptr) {
        cudaError_t e = cudaMemcpyAsync(ctrl_host, ctrl, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) {
            set_error("gpet_training_sets_f64: control block copy: %s", cudaGetErrorString(e));
            return GPET_ERR_CUDA;
        }
    }
    return GPET_OK;
}
