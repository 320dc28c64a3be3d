"""Host side of the "Render" step of Render-Localise-Lift, mirroring the reference's helpers by name:
`normalize_mesh`, `generate_sam_inp_objs` (utils/demo_utils.py:130-256), `get_rasterizer`,
`project_vertices_and_create_mask`, `render_mesh` (preprocess_data/render_mesh_utils.py:115-198).  The reference runs
pytorch3d (look_at_view_transform + FoVPerspectiveCameras + MeshRasterizer + HardPhongShader); here the camera set-up is
a few lines of float32 host arithmetic and the rasteriser / shader are the sm_100a kernels behind
`ivlm_rasterize_mesh` / `ivlm_shade_phong`.  A "mesh" is a `(verts [Nv,3], faces [Nf,3])` pair."""
from __future__ import annotations

import math
import os
from pathlib import Path

import numpy as np
import torch

from . import ops

FACES_PER_PIXEL = 1   # render_mesh_utils.py:24-25
BLUR_RADIUS = 0.0
LIGHT_LOCATIONS = [[0, 0, 3], [0, 0, 3], [0, 0, -3], [0, 0, -3]]   # utils/demo_utils.py:23-26
RENDER_IMG_SIZE = (1024, 1024)
YELLOW_VERTEX_COLOR = [1.00, 0.90, 0.30]
OBJECT_VIEWS_4 = {"frontleft": (1.5, 45, 315, 0.0, 0.0), "frontright": (1.5, 45, 45, 0.0, 0.0),   # demo_utils.py:196-201
                  "backleft": (1.5, 330, 135, 0.0, 0.0), "backright": (1.5, 330, 225, 0.0, 0.0)}


def normalize_mesh(vertices, scale_factor=1.0):
    """utils/demo_utils.py:130-145: centre on the centroid, scale the longest bounding-box side to `scale_factor`."""
    v = torch.as_tensor(vertices, dtype=torch.float32)
    c = v - v.mean(dim=0)
    size = (c.max(dim=0)[0] - c.min(dim=0)[0]).max()
    return c / size * scale_factor


def look_at_view_transform(dist, elev, azim):
    """pytorch3d's look_at_view_transform for one camera (degrees, looking at the origin, up = +Y) -> R, T, C (float32)."""
    f = torch.float32
    d, el, az = (torch.tensor(float(x), dtype=f) for x in (dist, elev, azim))
    el, az = el * (math.pi / 180.0), az * (math.pi / 180.0)
    C = torch.stack([d * torch.cos(el) * torch.sin(az), d * torch.sin(el), d * torch.cos(el) * torch.cos(az)])
    unit = lambda t: t / t.norm().clamp_min(1e-5)
    up = torch.tensor([0.0, 1.0, 0.0], dtype=f)
    z = unit(-C)
    x = unit(torch.linalg.cross(up, z))
    y = unit(torch.linalg.cross(z, x))
    if bool(torch.isclose(x, torch.zeros(3), atol=5e-3).all()):
        x = unit(torch.linalg.cross(y, z))
    R = torch.stack([x, y, z], dim=1)          # axes as columns; X_view = X_world @ R + T
    T = -(R.t() @ C)
    return R, T, C


def camera(camera_params, fov=60.0, znear=1.0):
    """get_rasterizer's camera (render_mesh_utils.py:115-121): (distance, elevation, azimuth, x_trans, y_trans) ->
    dict(R, T, C, s, z_clip) for ops.rasterize_mesh (FoVPerspectiveCameras defaults: fov 60 deg, znear 1, aspect 1)."""
    dist, elev, azim, x_trans, y_trans = [float(c) for c in camera_params]
    R, T, C = look_at_view_transform(dist, elev, azim)
    T = T.clone()
    T[1] += y_trans
    T[0] += x_trans
    # FoVPerspectiveCameras.compute_projection_matrix in float32: K00 = K11 = 2 znear / (max - min), max = tan(fov/2) znear
    half = torch.tan(torch.tensor(float(fov), dtype=torch.float32) * (math.pi / 180.0) / 2) * znear
    s = 2.0 * znear / (half - (-half))
    return dict(R=R.numpy(), T=T.numpy(), C=C.numpy(), s=np.float32(s.item()), z_clip=np.float32(znear / 2.0))


_CTX = {}


def _ctx(device=0):
    if not torch.cuda.is_available():
        raise RuntimeError("interactvlm_b200.render needs a CUDA device (sm_100a); there is no CPU fallback")
    idx = torch.device("cuda", device).index if not isinstance(device, int) else device
    if idx not in _CTX:
        _CTX[idx] = ops.Context(idx)
    return _CTX[idx]


def _mesh_tensors(mesh, device):
    verts, faces = mesh
    v = torch.as_tensor(verts, dtype=torch.float32).to(device).contiguous()
    f = torch.as_tensor(np.asarray(faces.cpu() if torch.is_tensor(faces) else faces).astype(np.int32)).to(device).contiguous()
    return v, f


def rasterize_views(mesh, camera_params_list, image_size=(512, 512), device=0, ctx=None, want_p2v=True):
    """All views of one mesh in one call (views are batched through the kernel grid) -> dict of CUDA tensors
    (pix_to_face [V,H,W], bary [V,H,W,3], p2v [V,H,W,3]) + the cameras."""
    ctx = ctx or _ctx(device)
    v, f = _mesh_tensors(mesh, ctx.device)
    cams = [camera(p) for p in camera_params_list]
    out = ops.rasterize_mesh(ctx, v, f, cams, int(image_size[0]), int(image_size[1]), want_p2v=want_p2v)
    out.update(cams=cams, verts=v, faces=f)
    return out


def project_vertices_and_create_mask(mesh, camera_params, contact_vertices, image_size=(512, 512), min_vertices=3,
                                     device=0, ctx=None):
    """render_mesh_utils.py:123-174, same returns: (mask uint8 [H,W] in {0,255}, pixel_to_vertices_map int64 [H,W,3]
    with -1 background, bary_coords float32 [H,W,3])."""
    r = rasterize_views(mesh, [camera_params], image_size, device, ctx)
    faces = r["faces"].long()
    hot = torch.zeros(r["verts"].shape[0], dtype=torch.bool, device=faces.device)
    cv = torch.as_tensor(sorted(set(int(c) for c in contact_vertices)), dtype=torch.long, device=faces.device)
    if cv.numel():
        hot[cv] = True
    face_hot = hot[faces].sum(1) >= min_vertices
    pix = r["pix_to_face"][0].long()
    mask = ((pix >= 0) & face_hot[pix.clamp_min(0)]).to(torch.uint8) * 255
    return mask.cpu().numpy(), r["p2v"][0].cpu().numpy(), r["bary"][0].cpu().numpy()


def render_mesh(mesh, camera_params, light_location, image_size=(512, 512), vertex_colors=None, device=0, ctx=None):
    """render_mesh_utils.py:177-198: HardPhong render with one point light -> uint8 [H,W,3]."""
    ctx = ctx or _ctx(device)
    r = rasterize_views(mesh, [camera_params], image_size, device, ctx, want_p2v=False)
    col = (torch.full_like(r["verts"], 0.85) if vertex_colors is None
           else torch.as_tensor(vertex_colors, dtype=torch.float32).to(ctx.device).contiguous())
    rgb = ops.shade_phong(ctx, r["verts"], r["faces"], col, r["cams"], [light_location], r["pix_to_face"], r["bary"])
    return rgb[0].cpu().numpy()


def lift_map_from_mesh(mesh, camera_params_list, image_size=RENDER_IMG_SIZE, device=0, ctx=None):
    """Mesh + cameras -> ops.LiftMap directly (what lift2d_dict.pkl would hold, without the pickle round trip)."""
    ctx = ctx or _ctx(device)
    r = rasterize_views(mesh, camera_params_list, image_size, device, ctx)
    return ops.LiftMap(ctx, r["p2v"].cpu().numpy(), r["bary"].cpu().numpy(), int(r["verts"].shape[0]))


def load_obj(path):
    """Minimal Wavefront reader (v / f records, polygons fan-triangulated, negative indices) -> (verts f32, faces i64)."""
    vs, fs = [], []
    with open(path) as fh:
        for line in fh:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                vs.append([float(x) for x in p[1:4]])
            elif p[0] == "f":
                idx = [int(tok.split("/")[0]) for tok in p[1:]]
                idx = [i - 1 if i > 0 else len(vs) + i for i in idx]
                fs.extend([idx[0], idx[k], idx[k + 1]] for k in range(1, len(idx) - 1))
    return torch.tensor(vs, dtype=torch.float32), torch.tensor(fs, dtype=torch.int64)


def generate_sam_inp_objs(obj_mesh_f, image_size=RENDER_IMG_SIZE, device=0):
    """utils/demo_utils.py:171-256: renders the 4 object views (grey = yellow vertex colour, and xyz-coloured) into
    `<dir>/sam_inp_objs/obj_render_{grey,color}_{view}.png` and writes `lift2d_dict.pkl` (pixel_to_vertices_map,
    bary_coords_map, num_vertices) next to them.  Returns the directory."""
    import cv2
    import joblib

    out_dir = Path(os.path.dirname(obj_mesh_f)) / "sam_inp_objs"
    if out_dir.exists():
        return out_dir
    out_dir.mkdir(parents=True)
    ctx = _ctx(device)
    vertices, faces = load_obj(obj_mesh_f)
    vertices = normalize_mesh(vertices)
    names, params = list(OBJECT_VIEWS_4), list(OBJECT_VIEWS_4.values())
    r = rasterize_views((vertices, faces), params, image_size, device, ctx)
    for prefix in ("grey", "color"):
        if prefix == "grey":
            col = torch.tensor(YELLOW_VERTEX_COLOR).repeat(vertices.shape[0], 1)
        else:
            lo, hi = vertices.min(dim=0)[0], vertices.max(dim=0)[0]
            col = (vertices - lo) / (hi - lo)
        col = (col * 0.8 + 0.1).to(ctx.device).contiguous()
        rgb = ops.shade_phong(ctx, r["verts"], r["faces"], col, r["cams"], LIGHT_LOCATIONS, r["pix_to_face"], r["bary"]).cpu().numpy()
        for k, name in enumerate(names):
            cv2.imwrite(str(out_dir / f"obj_render_{prefix}_{name}.png"), cv2.cvtColor(rgb[k], cv2.COLOR_RGB2BGR))
    p2v, bary = r["p2v"].cpu().numpy(), r["bary"].cpu().numpy()
    joblib.dump({"pixel_to_vertices_map": [p2v[k] for k in range(len(names))],
                 "bary_coords_map": [bary[k] for k in range(len(names))],
                 "num_vertices": int(vertices.shape[0])}, out_dir / "lift2d_dict.pkl")
    return out_dir


# ------------------------------------------------------------------------------------------------ point clouds (LEMON / PIAD)
def point_camera(camera_params, fov=60.0, znear=1.0):
    """get_rasterizer of preprocess_data/utils_obj_pc.py:28-32: like `camera`, but only the y translation is applied."""
    dist, elev, azim, _x_trans, y_trans = [float(c) for c in camera_params]
    return camera((dist, elev, azim, 0.0, y_trans), fov, znear)


def get_dynamic_radius(points):
    """utils_obj_pc.py:18-26: 0.004 x the bounding-box diagonal of the cloud."""
    p = np.asarray(points.detach().cpu() if torch.is_tensor(points) else points, dtype=np.float32).reshape(-1, 3)
    return float(0.004 * np.linalg.norm(p.max(0) - p.min(0)))


def project_points_to_image(points, camera_params, dynamic_radius=True, fixed_radius=0.005, image_size=(512, 512), num_point2pixel=1,
                            device=0, ctx=None):
    """utils_obj_pc.py:88-113 for `num_point2pixel == 1` -> pixel_to_point_map int64 [H,W] (-1 background), the `mapping` array
    of the `p2pmap_*.npz` files ObjectPCAfford3DPredictor reads (components.py:309)."""
    if num_point2pixel != 1:
        raise NotImplementedError("only the single-point-per-pixel maps the released pipeline writes (num_point2pixel=1)")
    ctx = ctx or _ctx(device)
    pts = torch.as_tensor(points, dtype=torch.float32).to(ctx.device).reshape(-1, 3).contiguous()
    radius = get_dynamic_radius(pts) if dynamic_radius else fixed_radius
    p2p = ops.rasterize_points(ctx, pts, [point_camera(camera_params)], int(image_size[0]), int(image_size[1]), radius)
    return p2p[0].cpu().numpy()


def create_affordance_mask(points, afford_indices, camera_params, dynamic_radius=True, fixed_radius=0.005, image_size=(512, 512),
                           device=0, ctx=None):
    """utils_obj_pc.py:115-137 (single point per pixel): 255 where the visible point carries the affordance."""
    p2p = project_points_to_image(points, camera_params, dynamic_radius, fixed_radius, image_size, 1, device, ctx)
    mask = np.zeros(tuple(image_size), np.uint8)
    mask[np.isin(p2p, list(set(int(i) for i in np.asarray(afford_indices).reshape(-1))))] = 255
    return mask, p2p
