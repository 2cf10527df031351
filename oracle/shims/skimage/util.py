"""TEST INFRASTRUCTURE ONLY. Stand-in for skimage.util.random_noise(mode='gaussian').

skimage >= 0.19 semantics restated: `rng = np.random.default_rng(seed)`,
`out = image + rng.normal(mean, sqrt(var), image.shape)`, clipped to [0, 1] for an
unsigned-range float image (all reference inputs are >= 0). skimage itself is not
installed, so the noise realisation is pinned to *this* stand-in, not to skimage."""
import numpy as np


def random_noise(image, mode="gaussian", seed=None, clip=True, mean=0.0, var=0.01, **kw):
    if mode != "gaussian":
        raise NotImplementedError("stand-in implements mode='gaussian' only")
    image = np.asarray(image, dtype=np.float64)
    low_clip = -1.0 if image.min() < 0 else 0.0
    rng = np.random.default_rng(seed)
    out = image + rng.normal(mean, var ** 0.5, image.shape)
    if clip:
        out = np.clip(out, low_clip, 1.0)
    return out
