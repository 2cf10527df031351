// Bench-input generator and trace-quality metrics on the device.
//
// Reference seams: gpet_utils.py:163-253 (construct_test_img: dark-above / bright-below step along one or two edge
// curves, optional occluding gaps, additive Gaussian noise clipped to [0, 1]) and :256-313 (trace_MSE, trace_relarea,
// trace_dicecoef).  The edge rows rint(A sin(N curvature x_j)) + M/2 are evaluated on the host (N numbers per image:
// numpy's sin decides the rint ties) and the noise field is supplied by the caller, so this kernel is pure fill work:
// 8 bytes written (+ 8 read for the noise) per pixel, HBM bound.
#include "gpet_common.cuh"

namespace gpet {

// x in range(N)[s:e] with python's slice semantics (negative bounds count from the end; gpet_utils.py:244-248 writes
// img[:, N-100:N-90] also for N < 100, where that slice is empty or wraps)
__device__ __forceinline__ bool in_py_slice(int x, int s, int e, int N) {
    if (s < 0) s = max(s + N, 0);
    if (e < 0) e = max(e + N, 0);
    return x >= min(s, N) && x < min(e, N);
}

__global__ void __launch_bounds__(256)
test_img_kernel(const int32_t* __restrict__ rows, const int32_t* __restrict__ rows2, int M, int N, double intensity,
                int gaps, const double* __restrict__ noise, double noise_sd, double* __restrict__ img) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= N) return;
    const int r1 = rows[(size_t)b * N + x];
    const int r2 = rows2 ? rows2[(size_t)b * N + x] : M;
    bool gap = false;
    if (gaps)      // gpet_utils.py:244-248
        gap = in_py_slice(x, 20, 30, N) || in_py_slice(x, N / 2, N / 2 + 10, N) || in_py_slice(x, N - 100, N - 90, N) ||
              in_py_slice(x, N / 4, N / 4 + 20, N);
    const int y0 = blockIdx.y * 32;
    for (int y = y0; y < min(M, y0 + 32); ++y) {
        double v = 0.0;
        if (!gap) {
            // python slices img[r:M, j]: a negative r counts from the end (gpet_utils.py:201)
            const int a1 = r1 < 0 ? max(M + r1, 0) : r1, a2 = r2 < 0 ? max(M + r2, 0) : r2;
            if (y >= a1) v = intensity;
            if (rows2 && y >= a2) v = 1.0 - intensity;
        }
        const size_t p = ((size_t)b * M + y) * N + x;
        if (noise) v = fmin(fmax(v + noise_sd * noise[p], 0.0), 1.0);    // random_noise(mode='gaussian', clip=True)
        img[p] = v;
    }
}

// out[b] = (mean squared row error, relative area difference, Jaccard index) of edge_pred[b][n][2] (y, x) against the
// true rows; the masks of gpet_utils.py:303-308 are column-wise half-open intervals [row, n), so their intersection and
// union have closed forms.  One warp per trace.
__global__ void __launch_bounds__(128)
trace_metrics_kernel(const int64_t* __restrict__ edge_pred, const int32_t* __restrict__ true_rows, int B, int n,
                     double* __restrict__ out) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    double se = 0.0, ta = 0.0, pa = 0.0, inter = 0.0, uni = 0.0;
    for (int j = lane; j < n; j += 32) {
        const long long p = edge_pred[((size_t)b * n + j) * 2];
        const long long t = true_rows[(size_t)b * n + j];
        const double d = (double)(p - t);
        se += d * d;
        ta += (double)(n - t);
        pa += (double)(n - p);
        const long long pc = p < 0 ? max(n + p, 0LL) : min(p, (long long)n), tc = t < 0 ? max(n + t, 0LL) : min(t, (long long)n);
        inter += (double)(n - max(pc, tc));
        uni += (double)(n - min(pc, tc));
    }
    se = warp_sum(se); ta = warp_sum(ta); pa = warp_sum(pa); inter = warp_sum(inter); uni = warp_sum(uni);
    if (lane == 0) {
        const double n2 = (double)n * (double)n;
        out[3 * b] = se / (double)n;
        out[3 * b + 1] = fabs((ta / n2 - pa / n2) / (ta / n2));
        out[3 * b + 2] = inter / uni;
    }
}

}  // namespace gpet

using namespace gpet;

extern "C" int gpet_test_img_f64(const int32_t* rows, const int32_t* rows2, int B, int M, int N, double intensity, int gaps,
                                 const double* noise, double noise_sd, double* img, void* stream) {
    GPET_REQUIRE(rows && img && B > 0 && M > 0 && N > 0, "gpet_test_img_f64: bad argument");
    GPET_SUPPORTED(B <= 65535, "gpet_test_img_f64: B too large for one launch");
    dim3 grid((N + 255) / 256, (M + 31) / 32, B);
    test_img_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rows, rows2, M, N, intensity, gaps, noise, noise_sd, img);
    return check_launch("test_img_kernel");
}

extern "C" int gpet_trace_metrics_f64(const int64_t* edge_pred, const int32_t* true_rows, int B, int n, double* out,
                                      void* stream) {
    GPET_REQUIRE(edge_pred && true_rows && out && B > 0 && n > 0, "gpet_trace_metrics_f64: bad argument");
    trace_metrics_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(edge_pred, true_rows, B, n, out);
    return check_launch("trace_metrics_kernel");
}
