"""Host side of the GPU image resize: coefficient tables of Pillow's antialiased separable resampling (8 bits per channel).

The reference resizes on the CPU with Pillow -- SAM views via `ResizeLongestSide.apply_image` (torchvision `resize` of a PIL
image, bilinear; model/segment_anything/utils/transforms.py:27-34, 102-113) and the CLIP image via `CLIPImageProcessor`
(bicubic, shortest edge 224, centre crop; run_demo.py:330-346).  Pillow's algorithm (libImaging/Resample.c, pinned here
by tests against the installed Pillow) is integer arithmetic once the per-output-pixel coefficient windows are known:
coefficients are doubles normalised to sum 1, converted to fixed point with 22 fractional bits, each pass accumulates
pixel * coeff from 2^21 and shifts, and the horizontal pass result is rounded to uint8 before the vertical pass.  The
tables are a few KB and are built here in float64 exactly as the C code does; the passes run in
`ivlm_resample_u8` (csrc/norm_elementwise.cu).  Bit-exact against Pillow.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bilinear(x: float) -> float:
    x = -x if x < 0 else x
    return 1.0 - x if x < 1.0 else 0.0


def _bicubic(x: float) -> float:
    a = -0.5
    x = -x if x < 0 else x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


FILTERS = {"bilinear": (_bilinear, 1.0), "bicubic": (_bicubic, 2.0)}


def precompute_coeffs(in_size: int, out_size: int, filt: str):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box.
    Returns (bounds [out,2] int32 = (first input index, count), coeffs [out, ksize] int32, ksize)."""
    fn, fsupport = FILTERS[filt]
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [fn((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        for x, w in enumerate(k):
            kk[xx, x] = int(-0.5 + w * (1 << PRECISION_BITS)) if w < 0 else int(0.5 + w * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def sam_target_size(h: int, w: int, long_side: int = 1024):
    """ResizeLongestSide.get_preprocess_shape (transforms.py:102-113)."""
    scale = long_side * 1.0 / max(h, w)
    return int(h * scale + 0.5), int(w * scale + 0.5)


def clip_target_size(h: int, w: int, short_side: int = 224):
    """CLIPImageProcessor resize: shortest edge -> 224, the other side int(224 * long / short) (transformers 4.31
    image_transforms.get_resize_output_image_size, default_to_square=False)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = short_side, int(short_side * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)
