"""Generates tests/golden/components.npz from the UNMODIFIED reference: the camera-pose encoders and the AttentionSplitter
(model/components.py:155-193,491-572) and `process_embeddings` (model/InteractVLM.py:268-294) are taken out of their files
with `ast` and executed as they are, for every camera-encoder type and every [SEG] / [HSEG] / [OSEG] token, in float32.
Run in the build container only:  python -m oracle.make_goldens_components"""
import ast
import textwrap
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = Path("/root/reference/model")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden" / "components.npz"
SEG, HSEG, OSEG = 320, 323, 324
V = 4


def _classes():
    ns = {"torch": torch, "nn": nn, "F": F}
    src = (REF / "components.py").read_text()
    for n in ast.parse(src).body:
        if isinstance(n, ast.ClassDef) and n.name in ("AttentionSplitter", "CamPoseEncoder", "ViewIndexCamPoseEncoder", "VIv1CamPoseEncoder"):
            exec(textwrap.dedent(ast.get_source_segment(src, n)), ns)
    src = (REF / "InteractVLM.py").read_text()
    cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "InteractVLMForCausalLM")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "process_embeddings")
    exec(textwrap.dedent(ast.get_source_segment(src, fn)), ns)
    return ns


def inputs():
    g = torch.Generator().manual_seed(21)
    emb = torch.randn(1, 256, generator=g)
    cam = torch.tensor([[.2, .125, .875, .5, .5], [.2, .875, .875, .5, .65], [.2, .125, .375, .5, .5], [.2, .875, .375, .5, .65]])
    return emb, cam


SHAPES = {
    "splitter": {"input_proj": (128, 256), "query_human": (128, 128), "query_object": (128, 128), "key": (128, 128),
                 "value": (128, 128), "output_proj": (256, 128)},
    "simple": {"linear1": (256, 5)},
    "view_index": {"spatial_encoder.0": (256, 5), "spatial_encoder.2": (256, 256),
                   **{f"view_transforms.{v}": (256, 256) for v in range(V)}},
    "vi_v1": {"spatial_encoder.0": (128, 5), "spatial_encoder.2": (128, 128),
              **{f"view_transforms.{v}": (256, 128) for v in range(V)}},
}


def component_weights(kind):
    """Seeded weights (not the modules' own initialisation, so that tests can rebuild them without the reference):
    N(0, 1.5 / sqrt(fan_in)) matrices -- large enough that the splitter's softmax is not flat -- and N(0, 0.2) biases."""
    g = torch.Generator().manual_seed({"splitter": 7, "simple": 11, "view_index": 12, "vi_v1": 13}[kind])
    sd = {}
    for name, (o, i) in SHAPES[kind].items():
        sd[name + ".weight"] = torch.randn(o, i, generator=g) * (1.5 / i ** 0.5)
        sd[name + ".bias"] = torch.randn(o, generator=g) * 0.2
    return sd


def main():
    ns = _classes()
    out = {}
    emb, cam = inputs()
    splitter = ns["AttentionSplitter"]()
    splitter.load_state_dict(component_weights("splitter"))
    for cam_type, cls in (("simple", "CamPoseEncoder"), ("view_index", "ViewIndexCamPoseEncoder"), ("vi_v1", "VIv1CamPoseEncoder")):
        enc = ns[cls]() if cam_type == "simple" else ns[cls](num_views=V)
        enc.load_state_dict(component_weights(cam_type))
        for token_type in ("Gen", "Gen-Hu-Obj"):
            for tok_name, tok in (("seg", SEG), ("hseg", HSEG), ("oseg", OSEG)):

                class Self:
                    multiview_cam_cond = True
                    cam_encoder_type = cam_type
                    cam_pose_encoder = enc
                    multiview_channels = V
                    base_token_type = token_type
                    hseg_token_idx, oseg_token_idx = HSEG, OSEG
                    attention_splitter = splitter
                with torch.no_grad():
                    e = emb.unsqueeze(1).repeat(1, V, 1).clone()      # InteractVLM.py:585-588
                    r = ns["process_embeddings"](Self, e, cam, tok)
                out[f"out/{cam_type}/{token_type}/{tok_name}"] = r.numpy()
    np.savez_compressed(OUT, **out)
    print({k: v.shape for k, v in out.items() if k.startswith("out/")})


if __name__ == "__main__":
    main()
