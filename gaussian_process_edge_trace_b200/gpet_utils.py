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


def construct_test_img(size, amplitude, curvature, noise_level, ltype, intensity, gaps=False, noise_seed=1):
    """Synthetic edge image (reference gpet_utils.py:163-253; single-edge ltypes). The reference hard-codes the
    noise seed 1; `noise_seed` lets a bench draw many different images. Returns (img float64, edge_idx (N,2) [y,x])."""
    M, N = size
    img = np.zeros((M, N))
    x = np.linspace(-np.pi, np.pi, N)
    A = M // 2 if amplitude > M else amplitude // 2
    cols = np.arange(0, N, 1)
    if ltype == "sinusoidal":
        rows = (np.rint(A * np.sin(N * curvature * x)) + M // 2).astype("int")
    elif ltype == "co-sinusoidal":
        rows = (np.rint(A * np.cos(N * curvature * x)) + M // 2).astype("int")
    elif ltype == "straight":
        rows = np.full(N, M // 2, dtype=int)
    elif ltype == "diag":
        rows = cols.copy()
    else:
        raise NotImplementedError(f"ltype={ltype!r} (multi-edge test images are not part of the hot path)")
    below = np.arange(M)[:, None] >= rows[None, :]
    img[below] = intensity
    edge_idx = np.stack([rows, cols], axis=1)
    if gaps:
        img[:, 20:30] = 0
        img[:, N // 2:(N // 2 + 10)] = 0
        img[:, N - 100:N - 90] = 0
        img[:, N // 4:(N // 4 + 20)] = 0
    return gaussian_noise(img, noise_seed, 0.0, noise_level), edge_idx


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
