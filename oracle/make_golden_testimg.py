"""TEST INFRASTRUCTURE ONLY. tests/golden/testimg.npz: noise-free construct_test_img outputs and the three trace metrics of
the UNMODIFIED reference (gpet_utils.py:163-313), for every ltype. Run in the build container:
    OPENBLAS_NUM_THREADS=1 python oracle/make_golden_testimg.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402

_, gu, _ = ref_harness.load_reference()
out = {}
for k, (size, amp, curv, lt, inten, gaps) in enumerate([
        ((64, 64), 40, 2, "sinusoidal", 0.3, True), ((64, 64), 40, 3, "multi-sinusoidal", 0.3, False),
        ((96, 96), 70, 2, "close multi-sinusoidal", 0.25, True), ((48, 48), 200, 2, "co-sinusoidal", 0.4, False),
        ((40, 40), 10, 1, "diag", 0.5, False), ((40, 40), 10, 1, "straight", 0.5, True)]):
    img, edge = gu.construct_test_img(size, amp, curv, 0.0, lt, inten, gaps=gaps)
    out[f"c{k}_img"], out[f"c{k}_edge"] = img, np.asarray(edge)
    out[f"c{k}_args"] = np.array([size[0], size[1], amp, curv, inten, int(gaps)], dtype=np.float64)
    out[f"c{k}_ltype"] = lt
rng = np.random.default_rng(2)
N = 64
true = np.stack([np.clip(32 + np.rint(12 * np.sin(np.arange(N) / 7.0)), 0, N).astype(int), np.arange(N)], axis=1)
for k in range(4):
    pred = true.copy()
    pred[:, 0] = np.clip(true[:, 0] + rng.integers(-6, 7, size=N), 0, N)
    out[f"m{k}_pred"], out[f"m{k}_true"] = pred, true
    out[f"m{k}_vals"] = np.array([gu.trace_MSE(pred, true), gu.trace_relarea(pred, true), gu.trace_dicecoef(pred, true),
                                   gu.trace_dicecoef(pred, true, jaccard=True)])
np.savez_compressed(os.path.join(os.path.dirname(HERE), "tests", "golden", "testimg.npz"), **out)
print("written", len(out), "arrays")
