"""Driver of the pose refinement (reference: `python -m optim.fit`, optim/fit.py:60-315 + optim/data_io.py): reads what the
hot path wrote (`hcontact_vertices.npz`, `ocontact_vertices.npz`) next to the OSX / detection files, builds the initial object
pose (mask-centroid translation, normal-filtered contact pairs, contact ICP) and runs the Adam loop of `optim.py` on the sm_100a
kernels (soft silhouette, contact term, nearest neighbour).

Kept from the reference on purpose (they change the numbers a maintainer would compare):
  * `initial_translation` indexes ROW 1 of `torch.nonzero(mask)` -- the (row, col) of the second mask pixel -- where a column
    mean was probably intended (fit.py:124-128);
  * the object mesh is flipped in y and z after centring (data_io.py:188-190);
  * contact thresholds 0.5 (human) / 0.3 (object), normal filter [True, 90, -90] degrees, ICP 10 iterations without scale.
Not rebuilt: tensorboard / video logging, the Phong overlay of every iteration, `save_init`.  trimesh and pytorch3d are absent
offline: `mesh_centroid` (trimesh's area-weighted centroid) and `vertex_normals` (pytorch3d's area-weighted vertex normals)
follow the libraries' documented definitions ("parity unpinned" for those two helpers; the init / filter arithmetic is pinned
to goldens recorded from the reference's own lines, tests/golden/fit_init.npz).
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

from . import optim as PO

# optim/constants.py:5-8
OSX_FOCAL_VIRTUAL = (5000, 5000)
OSX_INPUT_BODY_SHAPE = (256, 192)
OSX_PRINCPT = (OSX_INPUT_BODY_SHAPE[1] / 2, OSX_INPUT_BODY_SHAPE[0] / 2)

# optim/utils.py:15-19
HUMAN_INFERENCE_FILE, OBJECT_MESH_FILE, OBJECT_DETECTION_FILE = "osx_human2.npz", "object_mesh.obj", "object_detection.json"


def default_options() -> dict:
    """optim/cfg/fit.yaml."""
    return {"init": {"translation_hum_centroid": True, "icp": {"run": True, "filter_contacts": [True, 90, -90], "est_scale": False,
                                                               "max_iter": 10}},
            "max_iter": 250, "vars": ["pose"], "early_stop": False,
            "loss_weights": {"mask_loss": {"w": 5.0, "kick_in": 0}, "centroid_loss": {"w": 1.0e-4, "kick_in": 0},
                             "contact_loss": {"w": 10.0, "kick_in": 0}}}


@dataclass
class HumanParams:      # optim/data_io.py:18-46
    vertices: torch.Tensor
    faces: torch.Tensor
    normals: torch.Tensor
    contact_verts: torch.Tensor
    centroid_offset: torch.Tensor
    bbox: torch.Tensor | None = None
    mask: torch.Tensor | None = None


@dataclass
class ObjectParams:     # optim/data_io.py:49-75
    vertices: torch.Tensor
    faces: torch.Tensor
    normals: torch.Tensor
    contact_verts: torch.Tensor
    mask: torch.Tensor
    scale: torch.Tensor
    bbox: torch.Tensor | None = None


@dataclass
class CameraParams:     # optim/data_io.py:78-93
    focal_length: torch.Tensor
    principal_point: torch.Tensor


def vertex_normals(verts: torch.Tensor, faces: torch.Tensor) -> torch.Tensor:
    """pytorch3d `Meshes.verts_normals_list()[0]` then unit length (data_io.py:40-41): every face adds its un-normalised normal
    (area weighting) to its three vertices."""
    f = faces.long()
    v0, v1, v2 = verts[f[:, 0]], verts[f[:, 1]], verts[f[:, 2]]
    n = torch.zeros_like(verts)
    n.index_add_(0, f[:, 1], torch.cross(v2 - v1, v0 - v1, dim=1))
    n.index_add_(0, f[:, 2], torch.cross(v0 - v2, v1 - v2, dim=1))
    n.index_add_(0, f[:, 0], torch.cross(v1 - v0, v2 - v0, dim=1))
    n = F.normalize(n, eps=1e-6, dim=1)
    return n / torch.norm(n, dim=1, keepdim=True).clamp_min(1e-12)


def mesh_centroid(verts: np.ndarray, faces: np.ndarray) -> np.ndarray:
    """trimesh `Trimesh.centroid`: mean of the triangle centroids weighted by triangle area."""
    tri = verts[faces]
    area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    if float(area.sum()) <= 0:
        return verts.mean(0)
    return (tri.mean(1) * area[:, None]).sum(0) / area.sum()


def get_camera_params(human_bbox, device="cuda") -> CameraParams:
    """optim/data_io.py:96-109: OSX's virtual camera scaled to the person box (x, y, w, h)."""
    b = torch.as_tensor(np.asarray(human_bbox), dtype=torch.float32, device=device)
    fv, bs, pt = OSX_FOCAL_VIRTUAL, OSX_INPUT_BODY_SHAPE, OSX_PRINCPT
    focal = torch.stack([fv[0] / bs[1] * b[2], fv[1] / bs[0] * b[3]])
    princpt = torch.stack([pt[0] / bs[1] * b[2] + b[0], pt[1] / bs[0] * b[3] + b[1]])
    return CameraParams(focal, princpt)


def _read_mask(path_png: Path, fallback):
    if path_png.exists():
        import cv2

        m = cv2.imread(str(path_png), cv2.IMREAD_GRAYSCALE)
        assert m is not None, f"Image not found at {path_png}"
        return m
    return np.array(fallback, dtype=np.uint8)


def load_params(human_inference_file, object_mesh_file, object_detection_file, device="cuda"):
    """optim/data_io.py:134-212 -> (HumanParams, ObjectParams, CameraParams) on `device`."""
    from .render import load_obj

    hfile = Path(human_inference_file)
    root = hfile.parent
    hz = np.load(hfile, allow_pickle=True)
    hv = np.asarray(hz["smpl_vertices"][0], np.float32)
    hf = np.asarray(hz["smpl_faces"]).astype(np.int64)
    hum_contacts = np.load(root / "hcontact_vertices.npz")["pred_contact_3d_smplx"]
    hmask_fallback = None
    if not (root / "human_mask.png").exists():
        hmask_fallback = json.load(open(root / "human_detection.json"))["mask"]
    hmask = _read_mask(root / "human_mask.png", hmask_fallback)
    cam = get_camera_params(hz["bbox_2"][0], device)
    c_off = mesh_centroid(hv.astype(np.float64), hf)
    t = lambda a, dt=torch.float32: torch.as_tensor(np.asarray(a), dtype=dt, device=device)
    hverts = t(hv - c_off.astype(np.float32))
    human = HumanParams(vertices=hverts, faces=t(hf, torch.int64), normals=vertex_normals(hverts, t(hf, torch.int64)),
                        contact_verts=t(np.asarray(hum_contacts).reshape(-1)), centroid_offset=t(c_off), bbox=t(hz["bbox_2"][0]),
                        mask=t(hmask, torch.uint8))
    ov, of_ = load_obj(object_mesh_file)
    ov = np.asarray(ov, np.float64)
    ov = ov - mesh_centroid(ov, np.asarray(of_))
    ov[:, 1] *= -1
    ov[:, 2] *= -1
    det = json.load(open(object_detection_file))
    oroot = Path(object_detection_file).parent
    oc = np.load(oroot / "ocontact_vertices.npz")["pred_contact_3d"]
    if oc.ndim == 2:
        oc = oc.squeeze(0)
    omask = _read_mask(oroot / "object_mask.png", None if (oroot / "object_mask.png").exists() else np.array(det["mask"]).squeeze())
    overts = t(ov)
    obj = ObjectParams(vertices=overts, faces=t(of_, torch.int64), normals=vertex_normals(overts, t(of_, torch.int64)),
                       contact_verts=t(oc), mask=t(omask, torch.uint8), scale=t(np.array([1.0])), bbox=t(np.array(det["bbox"])))
    return human, obj, cam


def initial_translation(obj_mask, h_verts, h_contact_mask, cam: CameraParams):
    """fit.py:120-137 (`translation_hum_centroid`): depth = mean z of the human contact vertices, x / y from the mask pixel the
    reference picks (see the module docstring), back-projected with the pinhole camera."""
    idx = torch.nonzero(obj_mask).float()
    z = h_verts[h_contact_mask, 2].mean()
    cx = idx[1].mean() - cam.principal_point[0]
    cy = idx[0].mean() - cam.principal_point[1]
    return torch.stack([cx * z / cam.focal_length[0], cy * z / cam.focal_length[1], z]).to(h_verts.device)


def filter_contacts_by_normals(h_norms, h_contact_mask, o_norms, o_contact_probs, filter_contacts):
    """fit.py:141-167: an object contact vertex survives when its normal makes less than filter[1] degrees (or, with a third
    entry, more than filter[2] degrees measured as cos < cos(filter[2])) with the reversed normal of at least one human contact
    vertex.  Returns (o_contact_mask, filtered contact probabilities) like the reference leaves them."""
    o_contact_mask = o_contact_probs > 0.3
    if not filter_contacts[0]:
        return o_contact_mask, o_contact_probs
    hn = F.normalize(-h_norms[h_contact_mask], p=2, dim=-1)
    on = F.normalize(o_norms[o_contact_mask], p=2, dim=-1)
    thr = torch.cos(torch.deg2rad(torch.tensor(float(filter_contacts[1]), dtype=torch.float32)))
    dots = torch.mm(on, hn.T)
    valid = dots > thr
    if len(filter_contacts) == 3:
        thr_neg = torch.cos(torch.deg2rad(torch.tensor(float(filter_contacts[2]), dtype=torch.float32)))
        valid = valid | (dots < thr_neg)
    best = valid.any(dim=1)
    o_contact_mask = o_contact_mask.clone()
    o_contact_mask[o_contact_mask.clone()] = best
    probs = o_contact_probs.clone()
    probs[~o_contact_mask] = 0.0
    return o_contact_mask, probs


@dataclass
class FitResult:
    rotation6d: torch.Tensor
    translation: torch.Tensor
    scale: torch.Tensor
    object_vertices: torch.Tensor
    history: list = field(default_factory=list)
    icp_rmse: float | None = None


def _setup_fit(human: HumanParams, obj: ObjectParams, cam: CameraParams, img_shape, opt: dict, ctx=None):
    """fit.py:86-215: initial pose (mask-centroid translation, normal-filtered contact ICP) -> (ObjPose_Opt, icp rmse)."""
    dev = human.vertices.device
    rot = PO.matrix_to_rot6d(torch.eye(3, device=dev)[None])[0]
    trans = torch.zeros(3, device=dev)
    scale = obj.scale.reshape(-1)[:1].clone()
    h_mask = human.contact_verts > 0.5
    o_probs = obj.contact_verts
    o_mask = o_probs > 0.3
    init, icp = opt["init"], opt["init"]["icp"]
    if init["translation_hum_centroid"]:
        trans = initial_translation(obj.mask, human.vertices, h_mask, cam)
    rmse = None
    if icp["run"]:
        o_mask, o_probs = filter_contacts_by_normals(human.normals, h_mask, obj.normals, o_probs, icp["filter_contacts"])
        o_mask = o_probs > 0.3                                         # extracted again since they may have changed (:173-175)
        if int(o_mask.sum()) >= 3 and int(h_mask.sum()) >= 3:
            sol = PO.ICP(obj.vertices[o_mask][None], human.vertices[h_mask][None], obj_contact_normals=obj.normals[o_mask][None],
                         hum_contact_normals=human.normals[h_mask][None], max_iterations=icp["max_iter"],
                         estimate_scale=icp["est_scale"],
                         init_transform=PO.SimilarityTransform(PO.rot6d_to_matrix(rot[None]), trans[None], scale))
            rot, trans = PO.matrix_to_rot6d(sol.RTs.R)[0], sol.RTs.T.reshape(3)
            rmse = float(sol.rmse) if sol.rmse is not None else None
    ren = PO.SSRenderer(tuple(img_shape[:2]), None, obj.faces, {"focal_length": cam.focal_length, "principal_point": cam.principal_point},
                        device=dev, ctx=ctx)
    model = PO.ObjPose_Opt(rot, trans, scale, {"vertices": human.vertices, "contact_verts": human.contact_verts,
                                               "centroid_offset": human.centroid_offset},
                           {"vertices": obj.vertices, "contact_verts": o_probs, "mask": obj.mask}, ren, vars=tuple(opt["vars"]), ctx=ctx).to(dev)
    return model, rmse


def _result(model, hist, rmse) -> FitResult:
    with torch.no_grad():
        verts = PO.apply_transformation(model.obj_vertices, model.rotation, model.translation, model.scale)
    return FitResult(model.rotation.detach(), model.translation.detach(), torch.as_tensor(model.scale).detach(), verts, hist, rmse)


def run_fit(human: HumanParams, obj: ObjectParams, cam: CameraParams, img_shape, opt: dict | None = None, record=False,
            ctx=None) -> FitResult:
    """fit.py:86-290 without the logging side: initial pose -> ObjPose_Opt -> Adam loop.  `record` keeps the per-iteration
    loss values (one host synchronisation per iteration, like the reference's progress bar); off for throughput runs."""
    opt = opt or default_options()
    model, rmse = _setup_fit(human, obj, cam, img_shape, opt, ctx)
    hist = PO.fit(model, opt["loss_weights"], max_iter=opt["max_iter"], early_stop=opt["early_stop"], record=record)
    return _result(model, hist, rmse)


_CTX_POOL: dict = {}


def run_fit_many(scenes, img_shape, opt: dict | None = None) -> list:
    """run_fit(record=False) for several (human, object, camera) scenes on one GPU, advanced together: every fit gets its own
    ivlm handle (scratch workspace) and stream, one iteration of each is captured as a CUDA graph and the replays are interleaved
    (optim.fit_many).  The reference runs one `python -m optim.fit` process per sample."""
    opt = opt or default_options()
    if opt["early_stop"]:
        return [run_fit(h, o, c, img_shape, opt) for h, o, c in scenes]
    from .ops import Context

    dev = scenes[0][0].vertices.device
    pool = _CTX_POOL.setdefault(dev.index or 0, [])
    while len(pool) < len(scenes):
        pool.append(Context(dev.index or 0))
    built = [_setup_fit(h, o, c, img_shape, opt, ctx=pool[i]) for i, (h, o, c) in enumerate(scenes)]
    PO.fit_many([m for m, _ in built], opt["loss_weights"], max_iter=opt["max_iter"])
    return [_result(m, [], rmse) for m, rmse in built]


def save_obj(path, verts, faces):
    """`final.obj` of fit.py:296-301 (one object)."""
    v, f = np.asarray(verts.detach().cpu()), np.asarray(torch.as_tensor(faces).cpu()).astype(np.int64) + 1
    with open(path, "w") as fh:
        for p in v:
            fh.write(f"v {p[0]:.6f} {p[1]:.6f} {p[2]:.6f}\n")
        for t in f:
            fh.write(f"f {t[0]} {t[1]} {t[2]}\n")


def main(input_path, opt: dict | None = None, out_root=None, device="cuda"):
    """`python -m optim.fit --input_path <image>` (fit.py:60-315): the sample folder holds the image next to osx_human2.npz,
    object_mesh.obj, object_detection.json and the two contact files.  Writes results/final.obj."""
    import cv2

    input_path = Path(input_path)
    out = (input_path / "results") if out_root is None else Path(out_root) / input_path.parts[-1]
    out.mkdir(exist_ok=True, parents=True)
    img = cv2.imread(str(input_path))
    assert img is not None, f"Image not found at {input_path}"
    d = input_path.parent
    human, obj, cam = load_params(d / HUMAN_INFERENCE_FILE, d / OBJECT_MESH_FILE, d / OBJECT_DETECTION_FILE, device)
    res = run_fit(human, obj, cam, img.shape, opt)
    save_obj(out / "final.obj", res.object_vertices + human.centroid_offset, obj.faces)
    return res
