"""Generates tests/golden/contact_loss.npz from the UNMODIFIED reference: the body of ObjPose_Opt.contact_loss is taken
from /root/reference/optim/optimizer.py with `ast` (the module itself cannot be imported here: tensorboard / pytorch3d /
matplotlib are absent) and executed as is, in float32 and float64, with autograd for the gradient.
Run in the build container only (needs /root/reference):  python -m oracle.make_goldens_optim"""
import ast
import textwrap
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/optim/optimizer.py")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden" / "contact_loss.npz"


def reference_contact_loss():
    src = REF.read_text()
    cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "ObjPose_Opt")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "contact_loss")
    code = textwrap.dedent(ast.get_source_segment(src, fn))
    ns = {"torch": torch}
    exec(code, ns)
    return ns["contact_loss"]


def inputs(seed, n_obj, n_hum):
    g = np.random.default_rng(seed)
    obj = g.normal(size=(n_obj, 3)) * 0.3 + np.array([0.2, 0.0, 0.5])
    hum = g.normal(size=(n_hum, 3)) * np.array([0.25, 0.6, 0.15])
    p = np.clip(g.beta(0.4, 1.5, size=n_obj), 0, 1)
    q = np.clip(g.beta(0.3, 2.0, size=n_hum), 0, 1)
    return obj.astype(np.float32), hum.astype(np.float32), p.astype(np.float32), q.astype(np.float32)


CASES = {"small": (0, 257, 300), "smplx": (1, 3000, 10475), "coincident": (2, 64, 64)}


def main():
    fn = reference_contact_loss()
    out = {}
    for name, (seed, n_obj, n_hum) in CASES.items():
        obj, hum, p, q = inputs(seed, n_obj, n_hum)
        if name == "coincident":
            hum[:16] = obj[:16]  # zero distances: cdist's subgradient is 0 there
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            class Self:
                object_contact_probs = torch.from_numpy(p).to(dt)
                human_contact_probs = torch.from_numpy(q).to(dt)
            o = torch.from_numpy(obj).to(dt).requires_grad_(True)
            loss = fn(Self, o, torch.from_numpy(hum).to(dt))
            loss.backward()
            out[f"{name}_{tag}_loss"] = loss.detach().numpy()
            out[f"{name}_{tag}_grad"] = o.grad.numpy()
    np.savez_compressed(OUT, **out)
    print({k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})


if __name__ == "__main__":
    main()
