"""GPU-box helper: where one pose-refinement iteration spends its time (CUDA events; synthetic scene of
examples/demo_synthetic_fit.py)."""
import importlib.util
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from interactvlm_b200 import optim as PO  # noqa: E402

spec = importlib.util.spec_from_file_location("demo_fit", str(ROOT / "examples" / "demo_synthetic_fit.py"))
demo = importlib.util.module_from_spec(spec)
spec.loader.exec_module(demo)


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


for size in (256, 512):
    human, obj, ren, gt_t, _ = demo.make_scene(size)
    verts = (obj["vertices"] + gt_t).clone().requires_grad_(True)
    faces = ren.o_faces.to(torch.int32)
    G = torch.randn(size, size, device="cuda")
    for _ in range(3):
        a, z = PO.soft_silhouette(verts, faces, ren.cam, (size, size))
        (a * G).sum().backward()
    torch.cuda.synchronize()
    n = 10
    e0 = ev()
    for _ in range(n):
        a, z = PO.soft_silhouette(verts, faces, ren.cam, (size, size))
    e1 = ev()
    for _ in range(n):
        a, z = PO.soft_silhouette(verts, faces, ren.cam, (size, size))
        (a * G).sum().backward()
    e2 = ev()
    ov = (obj["vertices"] + gt_t).clone().requires_grad_(True)
    for _ in range(n):
        PO.contact_loss(ov, human["vertices"], obj["contact_verts"], human["contact_verts"]).backward()
    e3 = ev()
    for _ in range(n):
        PO.calculate_centroid(a.detach())
    e4 = ev()
    torch.cuda.synchronize()
    nf = int((torch.ops.aten.sum(a.detach() > 0)).item())
    print(f"{size}^2, {faces.shape[0]} faces: soft silhouette forward {e0.elapsed_time(e1) / n:.2f} ms, forward+backward "
          f"{e1.elapsed_time(e2) / n:.2f} ms, contact loss fwd+bwd ({ov.shape[0]} x {human['vertices'].shape[0]}) "
          f"{e2.elapsed_time(e3) / n:.3f} ms, centroid {e3.elapsed_time(e4) / n:.3f} ms; covered pixels {nf}", flush=True)
model, hist, info = demo.run(iters=60, size=256)
torch.cuda.synchronize()
e0 = ev()
hist = PO.fit(model, demo.LOSS_WEIGHTS, max_iter=60)
e1 = ev()
torch.cuda.synchronize()
print(f"fit loop at 256^2: {e0.elapsed_time(e1) / 60:.2f} ms per iteration")

for graph in (False, True):
    e0 = ev()
    PO.fit(model, demo.LOSS_WEIGHTS, max_iter=60, record=False, graph=graph)
    e1 = ev()
    torch.cuda.synchronize()
    print(f"fit loop at 256^2 without host read-backs (record=False), CUDA-graph replays = {graph}: {e0.elapsed_time(e1) / 60:.2f} ms per iteration")
import time
from interactvlm_b200 import fit as FIT
from interactvlm_b200.bench_fit import _scene
h_, o_, c_ = _scene(256, 0, torch.device("cuda"))
opt = FIT.default_options()
for it in (0, 250):
    opt["max_iter"] = it
    torch.cuda.synchronize(); t0 = time.perf_counter()
    FIT.run_fit(h_, o_, c_, (256, 256), opt, record=False)
    torch.cuda.synchronize()
    print(f"run_fit with {it} iterations: {(time.perf_counter() - t0) * 1e3:.1f} ms wall", flush=True)
