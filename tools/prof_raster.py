"""GPU-box helper: time of the rasteriser (4 views x 1024^2, the reference's RENDER_IMG_SIZE) for meshes of growing size,
CUDA events around the whole ivlm_rasterize_mesh call (it contains one stream synchronisation), plus the Phong shader,
against the algorithmic output bytes (40 B per pixel: face id, 3 barycentrics, 3 int64 vertex ids).  (The CPU-oracle times quoted
in profiles/r1_raster.txt were taken with tests-only code: tools/ does not import oracle/.)"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from interactvlm_b200 import ops, render as R, synthetic as S  # noqa: E402

ctx = ops.Context(0)
views = list(R.OBJECT_VIEWS_4.values())
size = 1024
for n_lat, n_lon in ((24, 48), (96, 192), (256, 512), (512, 1024)):
    v, f = S.make_test_mesh("blob", n_lat=n_lat, n_lon=n_lon)
    v = R.normalize_mesh(torch.from_numpy(v)).numpy()
    r = R.rasterize_views((v, f), views, (size, size), ctx=ctx)
    vt, ft, cams = r["verts"], r["faces"], r["cams"]
    col = torch.full_like(vt, 0.8)
    for _ in range(3):
        ops.rasterize_mesh(ctx, vt, ft, cams, size, size)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    n = 10
    e0.record()
    for _ in range(n):
        out = ops.rasterize_mesh(ctx, vt, ft, cams, size, size)
    e1.record()
    for _ in range(n):
        ops.shade_phong(ctx, vt, ft, col, cams, R.LIGHT_LOCATIONS, out["pix_to_face"], out["bary"])
    e2.record()
    torch.cuda.synchronize()
    ms, ms_sh = e0.elapsed_time(e1) / n, e1.elapsed_time(e2) / n
    nbytes = 4 * size * size * 40
    line = f"{len(f):8d} faces: rasterise 4x{size}^2 {ms:7.3f} ms ({nbytes / ms / 1e6:7.1f} GB/s of output), phong {ms_sh:6.3f} ms, coverage {(out['pix_to_face'] >= 0).float().mean().item():.3f}"
    print(line, flush=True)
