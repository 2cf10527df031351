"""TEST INFRASTRUCTURE ONLY. `from skimage import restoration as rest` (gpet_utils.py:8); denoisers are off the hot path."""


def __getattr__(name):
    raise RuntimeError("scikit-image is not installed; denoisers are out of scope")
