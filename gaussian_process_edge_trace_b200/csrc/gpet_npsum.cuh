// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src: pairwise_sum), restated so that
// np.sum / np.mean / np.std of float64 vectors are reproduced bit for bit by one device thread.  The recursion of
// the original (halves rounded down to a multiple of 8, leaves of at most 128 elements) is run on an explicit stack:
// device recursion would need a stack-size limit that depends on n.
#pragma once

namespace gpet {

__device__ inline double np_pairwise_leaf(const double* a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
        r0 += a[i + 0]; r1 += a[i + 1]; r2 += a[i + 2]; r3 += a[i + 3];
        r4 += a[i + 4]; r5 += a[i + 5]; r6 += a[i + 6]; r7 += a[i + 7];
    }
    double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; ++i) res += a[i];
    return res;
}

__device__ inline double np_pairwise_sum(const double* a, int n) {
    if (n <= 128) return np_pairwise_leaf(a, n);
    int off[32], len[32], stage[32];
    double left[32];
    int sp = 0;
    off[0] = 0; len[0] = n; stage[0] = 0; left[0] = 0.0;
    double ret = 0.0;
    while (sp >= 0) {
        if (len[sp] <= 128) {
            ret = np_pairwise_leaf(a + off[sp], len[sp]);
            --sp;
            continue;
        }
        int n2 = len[sp] / 2;
        n2 -= n2 % 8;
        if (stage[sp] == 0) {
            stage[sp] = 1;
            off[sp + 1] = off[sp]; len[sp + 1] = n2; stage[sp + 1] = 0;
            ++sp;
        } else if (stage[sp] == 1) {
            left[sp] = ret;
            stage[sp] = 2;
            off[sp + 1] = off[sp] + n2; len[sp + 1] = len[sp] - n2; stage[sp + 1] = 0;
            ++sp;
        } else {
            ret = left[sp] + ret;
            --sp;
        }
    }
    return ret;
}

}  // namespace gpet
