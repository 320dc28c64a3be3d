"""GPU-box helper: stage drivers (ivlm_clip_encode / ivlm_seg_head / ivlm_mask_decode) against the op-level path, stage by stage."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from interactvlm_b200 import synthetic as S  # noqa: E402
from interactvlm_b200.config import IVLMConfig  # noqa: E402
from interactvlm_b200.model import InteractVLMForCausalLM  # noqa: E402
from interactvlm_b200.ops import Context  # noqa: E402
from oracle.make_goldens_model import TINY_SEED, tiny_inputs  # noqa: E402

ctx = Context(0)
cfg = IVLMConfig.tiny()
model = InteractVLMForCausalLM(cfg, S.make_state_dict(cfg, seed=TINY_SEED["weights"]), ctx=ctx)
ids, ans, clip, sam, cam = tiny_inputs(cfg, 2)
eng = model.eng
emb1 = eng.sam_encode(sam[0].cuda().bfloat16())
V = cfg.multiview_channels
emb = emb1.repeat(2, 1, 1)[: 2 * V].contiguous()
hid = torch.randn(2, cfg.hidden_size, generator=torch.Generator().manual_seed(5)).bfloat16().cuda()
res = {}
for stage in (True, False):
    eng.stage_abi = stage
    feats = eng.clip_encode(clip.cuda().bfloat16())
    prompt, e = eng.seg_prompt(hid, cam.cuda().bfloat16())
    res[stage] = [feats, prompt, e]
eng.stage_abi = True
prompt = res[False][1]
for stage in (True, False):
    eng.stage_abi = stage
    res[stage].append(eng.mask_decode(emb, prompt))
    res[stage].append(eng.mask_decode(emb[:V].contiguous(), prompt[:1].contiguous()))
for name, x, y in zip(("clip", "prompt", "emb", "lowres n=2", "lowres n=1"), res[True], res[False]):
    d = (x.float() - y.float()).abs()
    print(f"{name:>12}: shape {tuple(x.shape)}  max-abs diff {d.max().item():.5f}  (scale {y.float().abs().max().item():.3f})  equal {torch.equal(x, y)}")
    if name.startswith("lowres") and not torch.equal(x, y):
        per = d.flatten(1).max(1).values
        print("   per view:", [round(v, 4) for v in per.tolist()])
