"""Test infrastructure: make the UNMODIFIED reference (/root/reference) importable and runnable on CPU in the
build container (transformers 5.5 instead of the pinned 4.31, no deepspeed/pytorch3d, no ./data).

Nothing here edits or copies reference sources; the shims are monkey patches applied before import
(SURVEY.md section 8c lists them).  /root/reference does not exist on the GPU box, so only
oracle/make_goldens.py (run here, output committed under tests/golden/) and the optional
tests/test_reference_crosscheck.py (skipped when the path is absent) use this module.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

REFERENCE = Path(os.environ.get("IVLM_REFERENCE", "/root/reference"))


def available() -> bool:
    return (REFERENCE / "model" / "InteractVLM.py").exists()


_workdir = None


def workdir() -> Path:
    """Scratch cwd holding the synthetic ./data the reference reads with relative paths."""
    global _workdir
    if _workdir is None:
        _workdir = Path(tempfile.mkdtemp(prefix="ivlm_ref_"))
    return _workdir


def write_human_maps(p2v: np.ndarray, bary: np.ndarray, views=("topfront", "bottomfront", "topback", "bottomback")):
    """components.py:203-218 loads ./data/hcontact_vitruvian/{pixel_to_vertex_map_1024,bary_coords_map_1024}.npz."""
    d = workdir() / "data" / "hcontact_vitruvian"
    d.mkdir(parents=True, exist_ok=True)
    np.savez(d / "pixel_to_vertex_map_1024.npz", **{v: p2v[i] for i, v in enumerate(views)})
    np.savez(d / "bary_coords_map_1024.npz", **{v: bary[i] for i, v in enumerate(views)})


def tiny_clip_config(hidden=128, layers=3, heads=2, mlp=256, image=224, patch=14):
    from transformers import CLIPVisionConfig

    return CLIPVisionConfig(hidden_size=hidden, intermediate_size=mlp, num_hidden_layers=layers,
                            num_attention_heads=heads, image_size=image, patch_size=patch, hidden_act="quick_gelu",
                            layer_norm_eps=1e-5, projection_dim=hidden)


def apply(clip_config=None):
    """Apply the import shims and return the reference's `model.InteractVLM` module."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE}")
    import torch
    import transformers

    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    sys.dont_write_bytecode = True
    # 1. 'llava' is already a registered model type in transformers 5.x (llava_llama.py:166)
    transformers.AutoConfig.register = staticmethod(lambda *a, **k: None)
    transformers.AutoModelForCausalLM.register = staticmethod(lambda *a, **k: None)
    # 2. the MPT backbone (unused by InteractVLM) does not import on transformers 5.x
    stub = types.ModuleType("model.llava.model.language_model.llava_mpt")
    stub.LlavaMPTConfig = type("LlavaMPTConfig", (), {})
    stub.LlavaMPTForCausalLM = type("LlavaMPTForCausalLM", (), {})
    sys.modules["model.llava.model.language_model.llava_mpt"] = stub
    # 3. CLIP comes from the hub in the reference; build it locally from a config with random init
    cfg = clip_config or tiny_clip_config()
    transformers.CLIPVisionConfig.from_pretrained = classmethod(lambda cls, *a, **k: cfg)
    transformers.CLIPVisionModel.from_pretrained = classmethod(lambda cls, *a, **k: _clip_431_class()(_eager(cfg)))
    transformers.CLIPImageProcessor.from_pretrained = classmethod(lambda cls, *a, **k: transformers.CLIPImageProcessor())
    # 4. hard-coded .cuda() calls (InteractVLM.py:335,339,393,547,561) on a CPU-only box
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.cuda.empty_cache = lambda *a, **k: None
    # 5. relative ./data paths
    os.chdir(workdir())
    if str(REFERENCE) not in sys.path:
        sys.path.insert(0, str(REFERENCE))
    import importlib

    return importlib.import_module("model.InteractVLM")


def _clip_431_class():
    """transformers 5.x collects `hidden_states` with forward hooks that double-record once the tower has been called
    from inside another model's forward (probed: 7 entries instead of 4 after one LLaVA forward), which shifts
    `hidden_states[-2]`.  transformers 4.31 (the reference's pin) returns (embeddings, layer_1, ..., layer_L); this
    subclass builds exactly that tuple from the stock sub-modules."""
    import types as _t

    import transformers

    class CLIPVisionModel431(transformers.CLIPVisionModel):
        def forward(self, pixel_values=None, output_hidden_states=False, **kw):
            vm = self.vision_model
            h = vm.pre_layrnorm(vm.embeddings(pixel_values))
            states = [h]
            for layer in vm.encoder.layers:
                h = layer(h, None)
                h = h[0] if isinstance(h, tuple) else h
                states.append(h)
            return _t.SimpleNamespace(last_hidden_state=h, hidden_states=tuple(states))

    return CLIPVisionModel431


def _eager(cfg):
    cfg._attn_implementation = "eager"
    return cfg
