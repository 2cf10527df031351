"""TEST INFRASTRUCTURE ONLY. scikit-image stand-in (package absent in this image).
Only `util.random_noise` is functional; it is used by gpet_utils.construct_test_img
(gpet_utils.py:251) to add Gaussian noise to the synthetic bench image."""
from . import util, metrics, measure, restoration  # noqa: F401
