"""Oracle (TEST INFRASTRUCTURE): numpy restatement of Pillow's 8-bit antialiased resize (libImaging/Resample.c:
ImagingResampleHorizontal_8bpc / ImagingResampleVertical_8bpc), the arithmetic behind the reference's
`ResizeLongestSide.apply_image` (segment_anything/utils/transforms.py:27-34) and `CLIPImageProcessor` resize.
Pillow is a third-party dependency of the reference (installed here: 12.2); tests/test_resize_cpu.py pins this restatement
bit-for-bit against `PIL.Image.resize`, which is the golden source."""
import numpy as np

from interactvlm_b200.resample import PRECISION_BITS, precompute_coeffs


def _pass(img, bounds, kk, axis):
    """One separable pass along `axis` (1 = horizontal, 0 = vertical) on uint8 [H,W,C]."""
    src = np.moveaxis(img.astype(np.int64), axis, 0)           # [L, ...]
    out = np.empty((bounds.shape[0],) + src.shape[1:], np.int64)
    for o in range(bounds.shape[0]):
        x0, n = bounds[o]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for j in range(n):
            acc += src[x0 + j] * int(kk[o, j])
        out[o] = acc >> PRECISION_BITS
    return np.moveaxis(np.clip(out, 0, 255).astype(np.uint8), 0, axis)


def resize_u8(img: np.ndarray, out_h: int, out_w: int, filt: str) -> np.ndarray:
    """img uint8 [H,W,C] -> [out_h,out_w,C]; horizontal pass first, rounded to uint8, then vertical (Resample.c)."""
    H, W = img.shape[:2]
    x = img
    if out_w != W:
        b, k, _ = precompute_coeffs(W, out_w, filt)
        x = _pass(x, b, k, 1)
    if out_h != H:
        b, k, _ = precompute_coeffs(H, out_h, filt)
        x = _pass(x, b, k, 0)
    return x
