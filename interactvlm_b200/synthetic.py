"""Seeded synthetic data of the reference's shapes (no datasets or checkpoints exist offline).

Used by bench.py, the tests and oracle/make_goldens.py so that all of them see identical inputs.
Only numpy here (np.random.default_rng / PCG64 is reproducible across machines).
"""
from __future__ import annotations

import numpy as np

HUMAN_VIEWS = ("topfront", "bottomfront", "topback", "bottomback")  # preprocess_data/constants.py:315-382
N_SMPL = 6890
N_SMPLX = 10475
# normalize_cam_params(HUMAN_VIEW_DICT['4MV-Z_Vitru']['cam_params']) -- datasets/base_contact_dataset.py:37-50
HCONTACT_CAM_PARAMS = np.array([[.2, .125, .875, .5, .5], [.2, .875, .875, .5, .65],
                                [.2, .125, .375, .5, .5], [.2, .875, .375, .5, .65]], dtype=np.float32)


def _silhouette(rng, size: int, coverage: float) -> np.ndarray:
    """Union of ellipses covering roughly `coverage` of the image."""
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    mask = np.zeros((size, size), bool)
    target = coverage * size * size
    tries = 0
    while mask.sum() < target and tries < 64:
        cy, cx = rng.uniform(0.25, 0.75, 2) * size
        ry, rx = rng.uniform(0.05, 0.22, 2) * size
        mask |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
        tries += 1
    return mask


def make_mesh_lift_maps(n_verts: int = N_SMPL, n_views: int = 4, size: int = 1024, coverage: float = 0.18,
                        tile: int = 8, seed: int = 0):
    """pixel_to_vertex [V,size,size,3] int64 (-1 background) and bary [V,size,size,3] float32.

    Spatially coherent like a rasterised mesh: the image is cut into `tile`-pixel cells, every cell is two
    triangles over a lattice of vertices, neighbouring cells share vertices, barycentrics are the true
    coordinates of the pixel centre inside its triangle (SURVEY.md section 8d explains why i.i.d. random
    triangles would collapse every probability to ~0.5).
    """
    rng = np.random.default_rng(seed)
    p2v = np.full((n_views, size, size, 3), -1, np.int64)
    bary = np.zeros((n_views, size, size, 3), np.float32)
    nl = size // tile + 1
    yy, xx = np.mgrid[0:size, 0:size]
    ty, tx = yy // tile, xx // tile
    fy = ((yy % tile) + 0.5) / tile
    fx = ((xx % tile) + 0.5) / tile
    upper = (fx + fy) <= 1.0
    for v in range(n_views):
        sil = _silhouette(rng, size, coverage)
        lattice = rng.integers(0, n_verts, size=(nl, nl))
        v00, v01 = lattice[ty, tx], lattice[ty, tx + 1]
        v10, v11 = lattice[ty + 1, tx], lattice[ty + 1, tx + 1]
        # upper triangle (v00, v01, v10): weights (1-fx-fy, fx, fy); lower (v11, v10, v01): (fx+fy-1, 1-fx, 1-fy)
        a = np.where(upper, v00, v11)
        b = np.where(upper, v01, v10)
        c = np.where(upper, v10, v01)
        wa = np.where(upper, 1.0 - fx - fy, fx + fy - 1.0)
        wb = np.where(upper, fx, 1.0 - fx)
        wc = np.where(upper, fy, 1.0 - fy)
        tri = np.stack([a, b, c], -1)
        w = np.stack([wa, wb, wc], -1).astype(np.float32)
        p2v[v][sil] = tri[sil]
        bary[v][sil] = w[sil]
    return p2v, bary


def make_point_lift_maps(n_points: int = 2048, n_views: int = 4, size: int = 1024, coverage: float = 0.12,
                         seed: int = 0) -> np.ndarray:
    """pixel_to_point [V,size,size] int64, -1 background (p2pmap_*.npz['mapping'], components.py:309-327)."""
    rng = np.random.default_rng(seed + 1000)
    out = np.full((n_views, size, size), -1, np.int64)
    cell = 8
    nl = size // cell
    yy, xx = np.mgrid[0:size, 0:size]
    for v in range(n_views):
        sil = _silhouette(rng, size, coverage)
        lattice = rng.integers(0, n_points, size=(nl, nl))
        ids = lattice[yy // cell, xx // cell]
        out[v][sil] = ids[sil]
    return out


def make_mask_logits(batch: int, n_views: int = 4, size: int = 1024, seed: int = 0, std: float = 4.0) -> np.ndarray:
    """Smooth random logits [B,V,size,size] float32 (bilinear blow-up of a coarse random field)."""
    rng = np.random.default_rng(seed + 2000)
    coarse = rng.normal(0.0, std, size=(batch, n_views, size // 32 + 1, size // 32 + 1)).astype(np.float32)
    yi = np.linspace(0, coarse.shape[2] - 1.001, size, dtype=np.float32)
    y0 = yi.astype(np.int64)
    fy = (yi - y0)[None, None, :, None]
    rows = coarse[:, :, y0, :] * (1 - fy) + coarse[:, :, y0 + 1, :] * fy
    fx = (yi - y0)[None, None, None, :]
    out = rows[:, :, :, y0] * (1 - fx) + rows[:, :, :, y0 + 1] * fx
    return np.ascontiguousarray(out, dtype=np.float32)


def make_smplx_matrix(n_smplx: int = N_SMPLX, n_smpl: int = N_SMPL, seed: int = 0) -> np.ndarray:
    """Dense [n_smplx, n_smpl] barycentric-style mapping with 3 non-zeros per row summing to 1
    (the shape of SMPL_TO_SMPLX_MAPPING's 'matrix', utils/utils.py:428-443)."""
    rng = np.random.default_rng(seed + 3000)
    m = np.zeros((n_smplx, n_smpl), np.float32)
    cols = rng.integers(0, n_smpl, size=(n_smplx, 3))
    w = rng.dirichlet((1.0, 1.0, 1.0), size=n_smplx).astype(np.float32)
    for k in range(3):
        np.add.at(m, (np.arange(n_smplx), cols[:, k]), w[:, k])
    return m
