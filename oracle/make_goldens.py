"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference under
oracle/ref_shim.py) on seeded synthetic inputs.  Run in the build container:

    python -m oracle.make_goldens [lift] [decoder] [tiny]

Inputs are regenerated from seeds by the tests (interactvlm_b200/synthetic.py, oracle/weights.py); only the
reference's OUTPUTS (sub-sampled where large) are committed.  Test infrastructure, not product code.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"

from interactvlm_b200 import synthetic as S  # noqa: E402
from oracle import ref_shim  # noqa: E402

LIFT_SEED = dict(maps=0, logits=1, points=2, obj=3)
OBJ_NVERTS = 5000


def gold_lift():
    """HumanContact3DPredictor / ObjectMeshContact3DPredictor / ObjectPCAfford3DPredictor / convert_contacts."""
    import joblib

    p2v, bary = S.make_mesh_lift_maps(seed=LIFT_SEED["maps"])
    ref_shim.write_human_maps(p2v, bary)
    ref_shim.apply()
    from model.components import (HumanContact3DPredictor, ObjectMeshContact3DPredictor,
                                  ObjectPCAfford3DPredictor)
    from utils.utils import convert_contacts

    out = {}
    logits = S.make_mask_logits(2, seed=LIFT_SEED["logits"])
    human = HumanContact3DPredictor("4MV-Z_Vitru", 4)
    out["human"] = human([torch.from_numpy(logits[b]) for b in range(2)]).numpy()

    # object mesh: lift2d_dict.pkl with per-view lists (components.py:392-424)
    op2v, obary = S.make_mesh_lift_maps(n_verts=OBJ_NVERTS, seed=LIFT_SEED["obj"], coverage=0.25)
    wd = ref_shim.workdir()
    pkl = wd / "lift2d_dict.pkl"
    joblib.dump({"pixel_to_vertices_map": [op2v[v] for v in range(4)], "bary_coords_map": [obary[v] for v in range(4)],
                 "num_vertices": OBJ_NVERTS}, pkl)
    obj = ObjectMeshContact3DPredictor("4MV-Z_HM", 4)
    out["object_mesh"] = obj([torch.from_numpy(logits[0])], ds_names=["ocontact"], lift2d_dict_path=str(pkl)).numpy()

    # point cloud: p2pmap_*.npz next to the mask paths (components.py:309)
    p2p = S.make_point_lift_maps(seed=LIFT_SEED["points"])
    paths = []
    for v in range(4):
        np.savez(wd / f"obj_p2pmap_{v}.npz", mapping=p2p[v])
        paths.append(str(wd / f"obj_mask_{v}.png"))
    pc = ObjectPCAfford3DPredictor("4MV-Z_HM", 4)
    heat = 1.0 / (1.0 + np.exp(-logits))  # HM view types feed sigmoid-ed maps (InteractVLM.py:452-456)
    out["points"] = pc([torch.from_numpy(heat[b]) for b in range(2)], ds_names=["oafford"] * 2,
                       mask_paths_list=[paths, paths]).numpy()

    mapping = S.make_smplx_matrix(seed=0)
    out["smplx"] = convert_contacts(torch.from_numpy(out["human"]), torch.from_numpy(mapping)).numpy()
    np.savez_compressed(GOLD / "lift.npz", **out)
    print("lift goldens:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    GOLD.mkdir(parents=True, exist_ok=True)
    which = sys.argv[1:] or ["lift", "decoder", "tiny"]
    torch.manual_seed(0)
    if "lift" in which:
        gold_lift()
    if "decoder" in which or "tiny" in which:
        from oracle import make_goldens_model

        if "decoder" in which:
            make_goldens_model.gold_decoder()
        if "tiny" in which:
            make_goldens_model.gold_tiny()
