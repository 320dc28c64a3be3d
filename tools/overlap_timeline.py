"""Timeline of the overlapped step (SAM encoder chunks on the low-priority stream next to the decode steps):
CUDA-event timestamps relative to the fork, for a list of (sm_limit, limited_chunks, sam_chunk) settings."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench as Bn  # noqa: E402
from interactvlm_b200 import synthetic as S  # noqa: E402
from interactvlm_b200.config import IVLMConfig  # noqa: E402
from interactvlm_b200.model import InteractVLMForCausalLM  # noqa: E402

cfg = IVLMConfig.full()
dev = torch.device("cuda", 0)
sd = S.make_state_dict(cfg, seed=0, device=dev, gain=0.5)
model = InteractVLMForCausalLM(cfg, sd, device=0)
del sd
p2v, bary = S.make_mesh_lift_maps(seed=0)
model.set_human_lift_maps(p2v, bary)
B = 8
ids, ans, clip, sam, cam = Bn.make_batch(cfg, B, seed=1234)
clip, sam, cam = clip.to(dev), sam.to(dev), cam.to(dev)
sizes = [Bn.SIZE] * B


def step():
    return model.evaluate(clip, sam, ids, cam, sizes, sizes, max_new_tokens=Bn.N_ANS, scripted=ans)


def wall(n=3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for _ in range(2):
    step()
print(f"serial: {wall():.1f} ms/step")
settings = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(104, -1, 4)]
for lim, nlim, chunk in settings:
    model.enable_overlap(sm_limit=lim, limited_chunks=None if nlim < 0 else nlim, sam_chunk=chunk)
    for _ in range(2):
        step()
    ms = wall()
    model.overlap["trace"] = tr = []
    step()
    torch.cuda.synchronize()
    model.overlap["trace"] = None
    t0 = tr[0][1]
    print(f"sm_limit {lim} limited_chunks {nlim} sam_chunk {chunk}: {ms:.1f} ms/step | " +
          " ".join(f"{n}@{t0.elapsed_time(e):.0f}" for n, e in tr[1:]))
