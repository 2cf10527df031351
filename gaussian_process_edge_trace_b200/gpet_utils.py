"""Drop-in for the hot-path part of reference gp_edge_tracing/gpet_utils.py: kernel_builder (:10-61),
normalise (:65-91), comp_grad_img (:95-119, on the GPU), plus host helpers used by benches/reports
(construct_test_img :163-253, metrics :256-313). Denoisers and plotting are out of scope."""
import numpy as np
import torch

from ._cabi import GpetError, call, ptr


def kernel_builder(size, b2d=False, normalize=False, vertical_edges=False, unit=False):
    """Sobel-like edge filter (reference gpet_utils.py:10-61). Host side: the filter is tiny."""
    rows, cols = size
    mid_r, mid_c = rows // 2, cols // 2
    kernel = np.zeros(size)
    if unit:
        kernel[:mid_r, :] = 1
    else:
        dist_r = np.abs(np.arange(mid_r) - mid_r)[:, None]
        dist_c = np.abs(np.arange(cols) - mid_c)[None, :]
        kernel[:mid_r, :] = 1 + np.clip(mid_r + 1 - dist_r - dist_c, 0, None)
    kernel[mid_r + 1:, :] = -kernel[0:mid_r, :][::-1]
    if b2d:
        kernel = kernel[::-1].copy()
    if vertical_edges:
        kernel = kernel.T
    if normalize:
        kernel = kernel / kernel.max()
    return kernel


def normalise(img, minmax_val=(0, 1), astyp=np.float32):
    """float32 min-max normalisation (reference gpet_utils.py:65-91). Host helper for small arrays; the device
    paths use gpet_normalise_f32."""
    lo, hi = minmax_val
    out = np.asarray(img).astype(np.float32)
    out -= out.min()
    out /= out.max()
    out *= (hi - lo)
    out += lo
    return out.astype(astyp)


def comp_grad_img(img, kernel, norm=True, astyp=np.float32, device=None, return_tensor=False, exact=True):
    """Gradient image = convolve(img, kernel, edge-replicated) clipped at 0, float32 min-max normalised
    (reference gpet_utils.py:95-119; `norm` is ignored there too - the reference tests the function object).

    `img` may be [M, N] or a batch [B, M, N]; runs on the GPU through gpet_comp_grad_img_f64 (exact=True: fp64
    accumulation in scipy.ndimage's order, the float32 result is bit-identical to the reference) or
    gpet_comp_grad_img_fast_f32 (exact=False: float32 fused multiply-adds, within ~1e-6 of it - the north_star bar for the
    stencil is 1e-4 - and HBM bound instead of FP64-pipe bound)."""
    if not torch.cuda.is_available():
        raise GpetError("comp_grad_img needs a CUDA device (there is no CPU fallback)")
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    if torch.is_tensor(img):
        x = img.to(dev, dtype=torch.float64)
    else:
        a = np.asarray(img)
        if a.dtype not in (np.float64, np.float32):
            # scipy.ndimage keeps the input dtype (integer images wrap there); we promote instead - documented deviation
            a = a.astype(np.float64)
        x = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    single = x.ndim == 2
    if single:
        x = x[None]
    x = x.contiguous()
    B, M, N = x.shape
    k = torch.from_numpy(np.ascontiguousarray(kernel, dtype=np.float64)).to(dev)
    out = torch.empty((B, M, N), dtype=torch.float32, device=dev)
    mm = torch.empty((B, 2), dtype=torch.int32, device=dev)
    call("gpet_comp_grad_img_f64" if exact else "gpet_comp_grad_img_fast_f32", ptr(x), B, M, N, ptr(k), int(k.shape[0]), int(k.shape[1]), ptr(out), ptr(mm),
         torch.cuda.current_stream().cuda_stream)
    if single:
        out = out[0]
    if return_tensor:
        return out
    return out.cpu().numpy().astype(astyp)


def gaussian_noise(image, seed, mean=0.0, var=0.01):
    """Gaussian noise + clip to [0, 1]: what skimage.util.random_noise(mode='gaussian', seed=seed) does for a
    non-negative float image (reference gpet_utils.py:251; skimage itself is not a dependency here)."""
    rng = np.random.default_rng(seed)
    return np.clip(image + rng.normal(mean, var ** 0.5, image.shape), 0.0, 1.0)


def edge_rows(size, amplitude, curvature, ltype):
    """Row of the edge(s) in every column (reference gpet_utils.py:186-230): returns (rows int[N], rows2 int[N] | None);
    rows2 is the second edge of the multi-sinusoidal types (A//2 or A//6 below the first)."""
    M, N = size
    x = np.linspace(-np.pi, np.pi, N)
    A = M // 2 if amplitude > M else amplitude // 2
    rows2 = None
    if ltype in ("sinusoidal", "multi-sinusoidal", "close multi-sinusoidal"):
        rows = (np.rint(A * np.sin(N * curvature * x)) + M // 2).astype("int")
        if ltype == "multi-sinusoidal":
            rows2 = rows + A // 2
        elif ltype == "close multi-sinusoidal":
            rows2 = rows + A // 6
    elif ltype == "co-sinusoidal":
        rows = (np.rint(A * np.cos(N * curvature * x)) + M // 2).astype("int")
    elif ltype == "straight":
        rows = np.full(N, M // 2, dtype=int)
    elif ltype == "diag":
        rows = np.arange(N)
    else:
        raise ValueError(f"unknown ltype {ltype!r}")
    return rows, rows2


def construct_test_img(size, amplitude, curvature, noise_level, ltype, intensity, gaps=False, noise_seed=1):
    """Synthetic edge image (reference gpet_utils.py:163-253, every ltype). The reference hard-codes the noise seed 1;
    `noise_seed` lets a bench draw many different images. Returns (img float64, edge_idx [y, x]: N rows, or 2N for the
    two-edge types, first edge first)."""
    M, N = size
    rows, rows2 = edge_rows(size, amplitude, curvature, ltype)
    cols = np.arange(0, N, 1)
    img = np.zeros((M, N))
    for r, val in ((rows, intensity), (rows2, 1 - intensity)):
        if r is not None:
            start = np.where(r < 0, np.maximum(M + r, 0), r)          # img[r:M, j] with python's negative indices
            img[np.arange(M)[:, None] >= start[None, :]] = val
    edge_idx = np.stack([rows, cols], axis=1)
    if rows2 is not None:
        edge_idx = np.concatenate([edge_idx, np.stack([rows2, cols], axis=1)], axis=0)
    if gaps:
        img[:, 20:30] = 0
        img[:, N // 2:(N // 2 + 10)] = 0
        img[:, N - 100:N - 90] = 0
        img[:, N // 4:(N // 4 + 20)] = 0
    return gaussian_noise(img, noise_seed, 0.0, noise_level), edge_idx


def construct_test_img_batch(size, amplitudes, curvatures, noise_level, ltype, intensity, gaps=False, noise=None,
                             noise_seed=None, device=None):
    """B test images built on the device (gpet_test_img_f64): image b uses amplitudes[b] /
    curvatures[b]. noise: float64 device tensor [B, M, N] of standard normals, or None with noise_seed (torch's device
    generator: a different stream than numpy's default_rng - the images are bench inputs, not parity vectors), or both
    None for noise-free images. Returns (img float64 device tensor [B, M, N], rows int[B, N], rows2 int[B, N] | None)."""
    if not torch.cuda.is_available():
        raise GpetError("construct_test_img_batch needs a CUDA device")
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    M, N = size
    pairs = [edge_rows(size, a, c, ltype) for a, c in zip(amplitudes, curvatures)]
    rows = np.stack([p[0] for p in pairs]).astype(np.int32)
    rows2 = None if pairs[0][1] is None else np.stack([p[1] for p in pairs]).astype(np.int32)
    B = rows.shape[0]
    if noise is None and noise_seed is not None and noise_level > 0:
        gen = torch.Generator(device=dev)
        gen.manual_seed(int(noise_seed))
        noise = torch.randn((B, M, N), dtype=torch.float64, device=dev, generator=gen)
    d_rows = torch.from_numpy(rows).to(dev)
    d_rows2 = torch.from_numpy(rows2).to(dev) if rows2 is not None else None
    img = torch.empty((B, M, N), dtype=torch.float64, device=dev)
    call("gpet_test_img_f64", ptr(d_rows), ptr(d_rows2), B, M, N, float(intensity), int(bool(gaps)), ptr(noise),
         float(noise_level) ** 0.5, ptr(img), torch.cuda.current_stream().cuda_stream)
    return img, rows, rows2


def trace_metrics_batch(edges, true_rows, device=None):
    """trace_MSE / trace_relarea / trace_dicecoef (reference gpet_utils.py:256-313) of B traces at once on the device
    (gpet_trace_metrics_f64): edges int[B, n, 2] (y, x), true_rows int[B, n]. Returns dict of float64[B] arrays with the
    reference's roundings (4, 5, 4 decimals)."""
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    e = torch.as_tensor(np.ascontiguousarray(edges, dtype=np.int64)).to(dev)
    t_ = torch.as_tensor(np.ascontiguousarray(true_rows, dtype=np.int32)).to(dev)
    B, n = e.shape[0], e.shape[1]
    out = torch.empty((B, 3), dtype=torch.float64, device=dev)
    call("gpet_trace_metrics_f64", ptr(e), ptr(t_), B, n, ptr(out), torch.cuda.current_stream().cuda_stream)
    o = out.cpu().numpy()
    jacc = o[:, 2]
    return dict(mse=np.round(o[:, 0], 4), relarea=np.round(o[:, 1], 5), dice=np.round(2 * jacc / (jacc + 1), 4),
                jaccard=np.round(jacc, 4))


def trace_MSE(edge_pred, edge_true):
    """reference gpet_utils.py:256-269"""
    N = edge_pred.shape[0]
    return np.round((1 / N) * np.sum((edge_pred[:, 0] - edge_true[:, 0]) ** 2), 4)


def trace_relarea(edge_pred, edge_true):
    """reference gpet_utils.py:271-286"""
    N = edge_pred.shape[0]
    true_area = np.sum(N - edge_true[:, 0]) / N ** 2
    pred_area = np.sum(N - edge_pred[:, 0]) / N ** 2
    return np.round(np.abs((true_area - pred_area) / true_area), 5)


def trace_dicecoef(edge_pred, edge_true, jaccard=False):
    """reference gpet_utils.py:288-313"""
    N = edge_pred.shape[0]
    rows = np.arange(N)[:, None]
    pred_bin = (rows >= edge_pred[:, 0][None, :]).astype(np.float64)
    true_bin = (rows >= edge_true[:, 0][None, :]).astype(np.float64)
    jacc = np.sum(pred_bin * true_bin) / np.sum(np.clip(pred_bin + true_bin, 0, 1))
    return np.round(jacc, 4) if jaccard else np.round(2 * jacc / (jacc + 1), 4)
