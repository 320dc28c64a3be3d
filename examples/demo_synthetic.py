#!/usr/bin/env python
"""The hcontact flow of the reference's run_demo.py (:273-452) on the B200 path, with seeded synthetic weights and inputs
(no checkpoints or datasets exist offline): load -> prepare inputs on the GPU -> model.evaluate() -> SMPL->SMPL-X ->
`*_hcontact_vertices.npz`.  Needs a B200; `--config full` builds the 13B + ViT-H sized model (28 GB of random weights).

    python examples/demo_synthetic.py --out /tmp/ivlm_demo [--config tiny|full] [--batch 2]
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from interactvlm_b200 import harness as Hn  # noqa: E402
from interactvlm_b200 import synthetic as S  # noqa: E402
from interactvlm_b200.config import IVLMConfig  # noqa: E402
from interactvlm_b200.model import InteractVLMForCausalLM, save_pretrained  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="/tmp/ivlm_demo")
    ap.add_argument("--config", default="tiny", choices=["tiny", "full"])
    ap.add_argument("--batch", type=int, default=2)
    args = ap.parse_args()
    out = Path(args.out)
    out.mkdir(parents=True, exist_ok=True)
    cfg = IVLMConfig.tiny() if args.config == "tiny" else IVLMConfig.full()
    if args.config == "tiny":
        # exercise the checkpoint reader the way run_demo.py:134-136 does
        save_pretrained(out / "ckpt", cfg, S.make_state_dict(cfg, seed=0))
        model = InteractVLMForCausalLM.from_pretrained(out / "ckpt", torch_dtype=torch.bfloat16)
    else:
        model = InteractVLMForCausalLM(cfg, S.make_state_dict(cfg, seed=0, device="cuda:0", gain=0.5))
    model.get_model().initialize_vision_modules(model.get_model().config)
    model = model.bfloat16().cuda().eval()
    p2v, bary = S.make_mesh_lift_maps(seed=0)
    model.set_human_lift_maps(p2v, bary)          # real use: model.load_human_lift_maps("./data")
    to_smplx = Hn.ContactConverter(model, S.make_smplx_matrix(seed=0))

    rng = np.random.default_rng(1)
    B, V = args.batch, cfg.multiview_channels
    image_u8 = rng.integers(0, 256, (B, 224, 224, 3), dtype=np.uint8)             # CLIPImageProcessor output size
    views_u8 = rng.integers(0, 256, (B, V, 1024, 1024, 3), dtype=np.uint8)        # the four body renders
    ids, ans = S.make_prompt_ids(cfg, B, seed=2)
    cam = torch.from_numpy(np.broadcast_to(S.HCONTACT_CAM_PARAMS, (B, V, 5)).copy())
    clip, sam, resize_list = Hn.prepare_inputs(model, image_u8, views_u8)
    t0 = time.time()
    res = model.evaluate(clip, sam, torch.from_numpy(ids), cam, resize_list, original_size_list=resize_list,
                         contact_type="hcontact", max_new_tokens=ans.shape[1], scripted=torch.from_numpy(ans))
    torch.cuda.synchronize()
    print(f"evaluate: {time.time() - t0:.3f} s for {B} samples; output_ids {tuple(res['output_ids'].shape)}")
    for b in range(B):
        c = res["pred_contact_3d"][b:b + 1]
        f = Hn.save_hcontact(out / f"sample{b}", c, to_smplx(c))
        z = np.load(f)
        print(f"{f.name}: smplh {z['pred_contact_3d_smplh'].shape} smplx {z['pred_contact_3d_smplx'].shape} "
              f"contact vertices (>=0.5): {(z['pred_contact_3d_smplh'] >= 0.5).sum()}")


if __name__ == "__main__":
    main()
