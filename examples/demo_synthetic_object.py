#!/usr/bin/env python
"""The object flow of the reference's run_demo.py (oafford / ocontact, :205-270 and :454-463) on the B200 path:
object_mesh.obj -> generate_sam_inp_objs (GPU rasteriser + Phong shader instead of pytorch3d: 4 rendered views +
lift2d_dict.pkl) -> the rendered PNGs are the SAM inputs -> model.evaluate(contact_type='oafford',
lift2d_dict_path=...) -> per-vertex object contact -> `*_oafford_vertices.npz`.  Seeded synthetic weights and a
synthetic mesh (no checkpoints or datasets exist offline).  Needs a B200.

    python examples/demo_synthetic_object.py --out /tmp/ivlm_demo_obj [--config tiny|full]
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from interactvlm_b200 import harness as Hn  # noqa: E402
from interactvlm_b200 import render as R  # noqa: E402
from interactvlm_b200 import synthetic as S  # noqa: E402
from interactvlm_b200.config import IVLMConfig  # noqa: E402
from interactvlm_b200.model import InteractVLMForCausalLM  # noqa: E402


def write_obj(path, verts, faces):
    with open(path, "w") as fh:
        fh.writelines(f"v {x[0]:.7f} {x[1]:.7f} {x[2]:.7f}\n" for x in verts)
        fh.writelines(f"f {t[0] + 1} {t[1] + 1} {t[2] + 1}\n" for t in faces)


def run(out: Path, config: str = "tiny", model=None, image_size=R.RENDER_IMG_SIZE):
    import cv2

    out.mkdir(parents=True, exist_ok=True)
    cfg = IVLMConfig.tiny() if config == "tiny" else IVLMConfig.full()
    if model is None:
        model = InteractVLMForCausalLM(cfg, S.make_state_dict(cfg, seed=0, device="cuda:0", gain=0.5 if config == "full" else 1.0))
    sample = out / "sample_object"
    sample.mkdir(exist_ok=True)
    verts, faces = S.make_test_mesh("blob", n_lat=64, n_lon=128)
    write_obj(sample / "object_mesh.obj", verts * 3.0 + 0.7, faces)            # arbitrary units: normalize_mesh undoes it
    obj_dir = R.generate_sam_inp_objs(str(sample / "object_mesh.obj"), image_size=image_size)   # run_demo.py:230-234
    names = list(R.OBJECT_VIEWS_4)
    views = np.stack([cv2.cvtColor(cv2.imread(str(obj_dir / f"obj_render_color_{n}.png")), cv2.COLOR_BGR2RGB) for n in names])
    rng = np.random.default_rng(1)
    image_u8 = rng.integers(0, 256, (1, 224, 224, 3), dtype=np.uint8)
    ids, ans = S.make_prompt_ids(cfg, 1, seed=2)
    cam = torch.zeros((1, len(names), 5))                                       # object views carry no camera conditioning
    clip, sam, resize_list = Hn.prepare_inputs_from_raw(model, image_u8, views[None])
    res = model.evaluate(clip, sam, torch.from_numpy(ids), cam, resize_list, original_size_list=[tuple(image_size)],
                         lift2d_dict_path=str(obj_dir / "lift2d_dict.pkl"), contact_type="oafford",
                         max_new_tokens=ans.shape[1], scripted=torch.from_numpy(ans))
    contact = res["pred_contact_3d"]
    f = Hn.save_ocontact(sample / "sample_object", contact)
    return model, res, f, obj_dir


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="/tmp/ivlm_demo_obj")
    ap.add_argument("--config", default="tiny", choices=["tiny", "full"])
    args = ap.parse_args()
    _, res, f, obj_dir = run(Path(args.out), args.config)
    z = np.load(f)
    key = list(z.keys())[0]
    print(f"rendered views + lift maps: {sorted(p.name for p in obj_dir.iterdir())}")
    print(f"{f.name}: {key} {z[key].shape}, contact vertices (>= 0.5): {(z[key] >= 0.5).sum()}, "
          f"pred_masks {[tuple(m.shape) for m in res['pred_masks']]}")


if __name__ == "__main__":
    main()
