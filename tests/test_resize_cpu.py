"""CPU: the resize oracle (oracle/resize.py) and the host coefficient builder (interactvlm_b200/resample.py) against the
golden source -- Pillow itself, the library the reference resizes with -- bit for bit."""
import numpy as np
import pytest
from PIL import Image

from interactvlm_b200 import resample as R
from oracle import resize as OR


@pytest.mark.parametrize("h,w,oh,ow,filt", [
    (480, 640, 768, 1024, "bilinear"),     # SAM: upscale to longest side 1024
    (1365, 2048, 683, 1024, "bilinear"),   # SAM: antialiased downscale
    (480, 640, 224, 298, "bicubic"),       # CLIP: shortest edge 224
    (100, 37, 224, 82, "bicubic"),         # upscale, odd sizes
    (64, 64, 64, 17, "bilinear"),          # horizontal pass only
    (33, 64, 5, 64, "bicubic"),            # vertical pass only
])
def test_oracle_resize_is_bit_exact_vs_pillow(h, w, oh, ow, filt):
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR if filt == "bilinear" else Image.BICUBIC))
    got = OR.resize_u8(img, oh, ow, filt)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_target_sizes_match_reference_helpers():
    assert R.sam_target_size(480, 640) == (768, 1024)
    assert R.sam_target_size(1365, 2048) == (683, 1024)     # int(x + 0.5) rounding of transforms.py:110-113
    assert R.sam_target_size(1024, 1024) == (1024, 1024)
    assert R.clip_target_size(480, 640) == (224, 298)
    assert R.clip_target_size(640, 480) == (298, 224)


def test_coefficient_windows_are_normalised():
    for n_in, n_out, f in [(640, 1024, "bilinear"), (2048, 1024, "bilinear"), (640, 298, "bicubic")]:
        b, k, ks = R.precompute_coeffs(n_in, n_out, f)
        assert k.shape == (n_out, ks) and (b[:, 0] >= 0).all() and (b[:, 0] + b[:, 1] <= n_in).all()
        assert np.abs(k.sum(1) - (1 << R.PRECISION_BITS)).max() <= ks  # fixed-point rounding of a unit-sum window
