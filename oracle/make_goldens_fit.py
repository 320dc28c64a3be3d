"""Generates tests/golden/fit_init.npz from the UNMODIFIED reference: the initialisation block of optim/fit.py's main()
(lines "Shortened" .. just before the ICP call: contact thresholds, mask-centroid translation, normal filter of the contact
pairs) is cut out of the source text and executed as is on seeded inputs.  The module itself cannot be imported here
(pytorch3d, omegaconf, trimesh absent).  Run in the build container only:  python -m oracle.make_goldens_fit"""
import textwrap
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

REF = Path("/root/reference/optim/fit.py")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden" / "fit_init.npz"


def reference_block():
    lines = REF.read_text().splitlines()
    a = next(i for i, l in enumerate(lines) if l.strip() == "# Shortened")
    b = next(i for i, l in enumerate(lines) if "human_contact_pcd = Pointclouds" in l)
    return textwrap.dedent("\n".join(lines[a:b]))


def inputs(seed, nh=600, no=400, size=96):
    g = np.random.default_rng(seed)
    unit = lambda a: a / np.linalg.norm(a, axis=1, keepdims=True)
    hv = g.normal(size=(nh, 3)).astype(np.float32) * 0.4 + np.array([0, 0, 3.0], np.float32)
    ov = g.normal(size=(no, 3)).astype(np.float32) * 0.2
    hn = unit(g.normal(size=(nh, 3)) * 0.35 + np.array([0, 0, 1.0])).astype(np.float32)   # a cone around +z: the filter bites
    on = unit(g.normal(size=(no, 3))).astype(np.float32)
    hc, oc = g.beta(0.5, 0.8, nh).astype(np.float32), g.beta(0.6, 0.9, no).astype(np.float32)
    mask = np.zeros((size, size), np.uint8)
    mask[30:55, 40:70] = 1
    mask[10:12, 5:9] = 1
    return hv, hn, hc, ov, on, oc, mask, np.array([110.0, 105.0], np.float32), np.array([48.0, 47.0], np.float32)


def run_reference(seed, filter_contacts):
    hv, hn, hc, ov, on, oc, mask, focal, pp = inputs(seed)
    t = torch.from_numpy
    human_params = types.SimpleNamespace(vertices=t(hv), normals=t(hn), contact_verts=t(hc))
    object_params = types.SimpleNamespace(vertices=t(ov), normals=t(on), contact_verts=t(oc).clone(), mask=t(mask))
    camera_params = types.SimpleNamespace(focal_length=t(focal), principal_point=t(pp))
    icp = types.SimpleNamespace(run=True, filter_contacts=filter_contacts)
    opt = types.SimpleNamespace(init=types.SimpleNamespace(translation_hum_centroid=True, icp=icp))
    ns = dict(torch=torch, F=F, logging=types.SimpleNamespace(info=lambda *a, **k: None), human_params=human_params,
              object_params=object_params, camera_params=camera_params, opt=opt, print=lambda *a, **k: None,
              translation_init=torch.zeros(3))
    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        exec(reference_block(), ns)
    finally:
        torch.Tensor.cuda = saved
    return dict(translation=ns["translation_init"].numpy(), o_mask=ns["o_contact_mask"].numpy(),
                o_probs=object_params.contact_verts.numpy())


CASES = {"two_sided": (3, [True, 90, -90]), "one_sided": (4, [True, 60]), "off": (5, [False, 90, -90])}

if __name__ == "__main__":
    out = {}
    for name, (seed, fc) in CASES.items():
        for k, v in run_reference(seed, fc).items():
            out[f"{name}_{k}"] = v
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})
