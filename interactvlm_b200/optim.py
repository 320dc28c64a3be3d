"""Pose-refinement pieces (reference: optim/): the contact term of ObjPose_Opt (optim/optimizer.py:80-96) as one fused
sm_100a pass that returns the value and the gradient w.r.t. the object vertices, wrapped as a torch.autograd.Function so
that it drops into the reference's optimisation loop (`loss_dict["contact_loss"] = self.contact_loss(obj_vertices,
self.human_vertices)`, optimizer.py:137).  Human vertices and both probability vectors are buffers in the reference
(optimizer.py:52-66): no gradient is produced for them."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from .ops import Context, P

_CTX = {}


def _ctx(device) -> Context:
    if not torch.cuda.is_available():
        raise RuntimeError("interactvlm_b200.optim needs a CUDA device (sm_100a); there is no CPU fallback")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _CTX:
        _CTX[idx] = Context(idx)
    return _CTX[idx]


def _f32(t, name):
    if not (t.is_cuda and t.dtype == torch.float32):
        raise ValueError(f"{name}: expected a CUDA float32 tensor, got {t.dtype} on {t.device}")
    return t.contiguous()


class _ContactLoss(torch.autograd.Function):
    @staticmethod
    def forward(fctx, obj_verts, human_verts, obj_probs, human_probs, ctx):
        o, h = _f32(obj_verts, "obj_verts"), _f32(human_verts, "human_verts")
        p, q = _f32(obj_probs, "object_contact_probs"), _f32(human_probs, "human_contact_probs")
        if o.dim() != 2 or o.shape[1] != 3 or h.dim() != 2 or h.shape[1] != 3 or p.shape != o.shape[:1] or q.shape != h.shape[:1]:
            raise ValueError(f"contact_loss: shapes {tuple(o.shape)}, {tuple(h.shape)}, {tuple(p.shape)}, {tuple(q.shape)}")
        loss = torch.empty((1,), device=o.device, dtype=torch.float32)
        grad = torch.empty_like(o) if obj_verts.requires_grad else None
        L.check(ctx.lib.ivlm_contact_loss(ctx.h, P(o), P(p), P(h), P(q), C.c_int32(o.shape[0]), C.c_int32(h.shape[0]), P(loss),
                                          P(grad), ctx.stream), "contact_loss")
        fctx.save_for_backward(grad)
        return loss[0]

    @staticmethod
    def backward(fctx, g):
        (grad,) = fctx.saved_tensors
        return (grad * g if grad is not None else None), None, None, None, None


def contact_loss(obj_verts, human_verts, object_contact_probs, human_contact_probs, ctx: Context | None = None):
    """ObjPose_Opt.contact_loss (optim/optimizer.py:80-96): sum_ij p_i q_j |o_i - h_j| / (sum p * sum q) -> 0-dim tensor,
    differentiable w.r.t. `obj_verts`."""
    ctx = ctx or _ctx(obj_verts.device)
    return _ContactLoss.apply(obj_verts, human_verts, object_contact_probs, human_contact_probs, ctx)


# ------------------------------------------------------------------------------------------------ contact ICP
from typing import NamedTuple, List, Optional  # noqa: E402


class SimilarityTransform(NamedTuple):   # optim/icp/icp.py:24-27
    R: torch.Tensor
    T: torch.Tensor
    s: torch.Tensor


class ICPSolution(NamedTuple):           # optim/icp/icp.py:30-35
    converged: bool
    rmse: Optional[torch.Tensor]
    Xt: torch.Tensor
    RTs: SimilarityTransform
    t_history: List[SimilarityTransform]


def knn1(x, y, ctx: Context | None = None):
    """Index (int64 [n]) and squared distance (fp32 [n]) of the nearest row of y [m,D] for every row of x [n,D], D <= 8:
    pytorch3d `knn_points(..., K=1)` as the contact ICP uses it (optim/icp/icp.py:187-196)."""
    ctx = ctx or _ctx(x.device)
    x, y = _f32(x, "x"), _f32(y, "y")
    if x.dim() != 2 or y.dim() != 2 or x.shape[1] != y.shape[1]:
        raise ValueError(f"knn1: expected [n,D] and [m,D], got {tuple(x.shape)} and {tuple(y.shape)}")
    idx = torch.empty((x.shape[0],), device=x.device, dtype=torch.int32)
    d2 = torch.empty((x.shape[0],), device=x.device, dtype=torch.float32)
    L.check(ctx.lib.ivlm_knn1(ctx.h, P(x), P(y), C.c_int32(x.shape[0]), C.c_int32(y.shape[0]), C.c_int32(x.shape[1]), P(idx),
                              P(d2), ctx.stream), "knn1")
    return idx.long(), d2


def corresponding_points_alignment(X, Y, weights=None, estimate_scale=False, allow_reflection=False, eps=1e-9):
    """optim/icp/icp.py:274-417 (Umeyama) for batched [b,n,d] tensors: s X R + T ~ Y.  A handful of O(n) reductions and a
    d x d SVD -- host-side glue around the nearest-neighbour kernel, written with torch ops."""
    b, n, dim = X.shape
    w = X.new_ones(b, n) if weights is None else weights.to(X.dtype)
    wsum = w.sum(1).clamp(eps)
    Xmu = (X * w[..., None]).sum(1, keepdim=True) / wsum[:, None, None]
    Ymu = (Y * w[..., None]).sum(1, keepdim=True) / wsum[:, None, None]
    Xc, Yc = (X - Xmu) * w[..., None], (Y - Ymu) * w[..., None]
    cov = torch.bmm(Xc.transpose(2, 1), Yc) / wsum[:, None, None]
    U, S, Vh = torch.linalg.svd(cov)
    E = torch.eye(dim, dtype=X.dtype, device=X.device)[None].repeat(b, 1, 1)
    if not allow_reflection:
        E[:, -1, -1] = torch.det(torch.bmm(U, Vh))
    R = torch.bmm(torch.bmm(U, E), Vh)
    if estimate_scale:
        s = (torch.diagonal(E, dim1=1, dim2=2) * S).sum(1) / ((Xc * Xc).sum((1, 2)) / wsum).clamp(eps)
        T = Ymu[:, 0, :] - s[:, None] * torch.bmm(Xmu, R)[:, 0, :]
    else:
        T = Ymu[:, 0, :] - torch.bmm(Xmu, R)[:, 0, :]
        s = T.new_ones(b)
    return SimilarityTransform(R, T, s)


def _apply_similarity_transform(X, R, T, s):
    return s[:, None, None] * torch.bmm(X, R) + T[:, None, :]


def ICP(obj_contact_pcd, hum_contact_pcd, init_transform: Optional[SimilarityTransform] = None, max_iterations: int = 100,
        relative_rmse_thr: float = 1e-6, estimate_scale: bool = False, allow_reflection: bool = False, verbose: bool = False,
        obj_contact_normals=None, hum_contact_normals=None, min_scale: float = None, scale_penalty: float = 10.0,
        ctx: Context | None = None) -> ICPSolution:
    """Drop-in for optim/icp/icp.py:38-268 with [b,n,3] CUDA tensors (the reference's `Pointclouds` inputs hold one cloud
    each: pass `pcd.points_padded()`).  Same loop, same outputs; the per-iteration `knn_points` over points ++ normals is the
    `ivlm_knn1` kernel.  Like the reference, the query cloud is assembled once before the loop (icp.py:176-185)."""
    Xt0, Y = obj_contact_pcd.float(), hum_contact_pcd.float()
    if Xt0.dim() != 3 or Y.dim() != 3 or Xt0.shape[0] != Y.shape[0] or Xt0.shape[2] != Y.shape[2]:
        raise ValueError("Point sets X and Y have to have the same number of batches and data dimensions.")
    b, n, dim = Xt0.shape
    mask = Xt0.new_ones(b, n)
    obj_init = Xt0.clone()
    if init_transform is not None:
        R, T, s = init_transform
        if R.shape != (b, dim, dim) or T.shape != (b, dim) or s.shape != (b,):
            raise ValueError("The initial transformation init_transform has to be a named tuple SimilarityTransform with "
                             "elements (R, T, s) of shapes (minibatch, dim, dim), (minibatch, dim) and (minibatch,).")
        R, T, s = R.float(), T.float(), s.float()
        Xt = _apply_similarity_transform(Xt0, R, T, s)
    else:
        R = torch.eye(dim, device=Xt0.device)[None].repeat(b, 1, 1)
        T, s, Xt = Xt0.new_zeros((b, dim)), Xt0.new_ones(b), Xt0
    q = torch.cat([Xt, obj_contact_normals.float()], -1) if obj_contact_normals is not None else Xt
    t = torch.cat([Y, -hum_contact_normals.float()], -1) if hum_contact_normals is not None else Y
    q, t = q.contiguous(), t.contiguous()
    prev, rmse, converged, history = None, None, False, []
    for iteration in range(max_iterations):
        nn = torch.stack([t[k][knn1(q[k], t[k], ctx)[0]] for k in range(b)])
        nn_pts, nn_normals = nn[..., :3], -nn[..., 3:]
        R, T, s = corresponding_points_alignment(obj_init, nn_pts, mask, estimate_scale, allow_reflection)
        Xt = _apply_similarity_transform(obj_init, R, T, s)
        history.append(SimilarityTransform(R, T, s))
        sq = ((Xt - nn_pts) ** 2).sum(2)
        rmse = ((sq * mask).sum(1) / mask.sum(1).clamp(1e-9)).sqrt()
        combined = rmse
        if nn_normals.shape[-1]:
            nt = _apply_similarity_transform(nn_normals, R, torch.zeros_like(T), s)
            combined = rmse + (1 - (nt * nn_normals).sum(2))          # [b,n] + [b] broadcast exactly as icp.py:219-224 does
        if min_scale is not None:
            combined = combined + scale_penalty * torch.clamp_min(s - min_scale, 0)
        rel = torch.ones_like(combined) if prev is None else (combined - prev) / prev
        if verbose:
            print(f"ICP iteration {iteration}: mean/max rmse = {rmse.mean():1.2e}/{rmse.max():1.2e}; mean relative rmse = {rel.mean():1.2e}")
        if bool((rel <= relative_rmse_thr).all()):
            converged = True
            break
        prev = combined
    return ICPSolution(converged, rmse, Xt, SimilarityTransform(R, T, s), history)


# ------------------------------------------------------------------------------------------------ soft silhouette renderer
import math  # noqa: E402

import numpy as np  # noqa: E402

from . import ops as _ops  # noqa: E402


def perspective_camera(focal_length, principal_point, image_size):
    """The camera of optim/renderer.py:29-46 -- pytorch3d PerspectiveCameras(in_ndc=False, R = diag(-1,-1,1), T = 0) -- as
    the dict ops.rasterize_mesh / soft_silhouette take: pixel focal lengths and principal point are converted to NDC with
    s = min(H, W) / 2 (f / s, -(c - size / 2) / s); no znear, so no z clipping."""
    H, W = int(image_size[0]), int(image_size[1])
    fl = torch.as_tensor(focal_length, dtype=torch.float64).reshape(-1)
    fl = fl.repeat(2) if fl.numel() == 1 else fl
    pp = torch.as_tensor(principal_point, dtype=torch.float64).reshape(2)
    s = min(H, W) / 2.0
    return dict(R=np.diag([-1.0, -1.0, 1.0]).astype(np.float32), T=np.zeros(3, np.float32), C=np.zeros(3, np.float32),
                fx=float(fl[0]) / s, fy=float(fl[1]) / s, cx=-(float(pp[0]) - W / 2.0) / s, cy=-(float(pp[1]) - H / 2.0) / s,
                z_clip=0.0)


class _SoftSilhouette(torch.autograd.Function):
    @staticmethod
    def forward(fctx, verts, faces, cam, H, W, sigma, blur_radius, K, ctx):
        v, f = _f32(verts, "verts"), faces.contiguous()
        if f.dtype != torch.int32 or not f.is_cuda or f.dim() != 2 or f.shape[1] != 3 or v.dim() != 2 or v.shape[1] != 3:
            raise ValueError("soft_silhouette: verts [Nv,3] float32 and faces [Nf,3] int32 CUDA tensors expected")
        dev = v.device
        alpha = torch.empty((H, W), device=dev, dtype=torch.float32)
        zbuf0 = torch.empty((H, W), device=dev, dtype=torch.float32)
        n_frag = torch.empty((H, W), device=dev, dtype=torch.int32)
        frag_face = torch.empty((K, H, W), device=dev, dtype=torch.int32)
        frag_sd = torch.empty((K, H, W), device=dev, dtype=torch.float32)
        frag_z = torch.empty((K, H, W), device=dev, dtype=torch.float32)
        cams = _ops._cam_array([cam])
        L.check(ctx.lib.ivlm_soft_silhouette(ctx.h, P(v), P(f), C.c_int32(v.shape[0]), C.c_int32(f.shape[0]), cams, C.c_int32(H),
                                             C.c_int32(W), C.c_float(sigma), C.c_float(blur_radius), C.c_int32(K), P(alpha), P(zbuf0),
                                             P(n_frag), P(frag_face), P(frag_sd), P(frag_z), ctx.stream), "soft_silhouette")
        del frag_z
        fctx.save_for_backward(v, f, n_frag, frag_face, frag_sd)
        fctx.meta = (cam, H, W, sigma, ctx)
        fctx.mark_non_differentiable(zbuf0)
        return alpha, zbuf0

    @staticmethod
    def backward(fctx, g_alpha, _g_z):
        v, f, n_frag, frag_face, frag_sd = fctx.saved_tensors
        cam, H, W, sigma, ctx = fctx.meta
        grad = torch.empty_like(v)
        L.check(ctx.lib.ivlm_soft_silhouette_backward(ctx.h, P(v), P(f), C.c_int32(v.shape[0]), C.c_int32(f.shape[0]),
                                                      _ops._cam_array([cam]), C.c_int32(H), C.c_int32(W), C.c_float(sigma),
                                                      P(g_alpha.float().contiguous()), P(n_frag), P(frag_face), P(frag_sd), P(grad),
                                                      ctx.stream), "soft_silhouette_backward")
        return grad, None, None, None, None, None, None, None, None


def soft_silhouette(verts, faces, cam, image_size, sigma=1e-4, blur_radius=None, faces_per_pixel=100, ctx: Context | None = None):
    """-> (alpha [H,W] differentiable w.r.t. verts, zbuf0 [H,W]).  pytorch3d MeshRasterizer(blur_radius, faces_per_pixel) +
    SoftSilhouetteShader(BlendParams(sigma)) for one camera (include/ivlm_b200.h: ivlm_soft_silhouette)."""
    ctx = ctx or _ctx(verts.device)
    if blur_radius is None:
        blur_radius = math.log(1.0 / 1e-4 - 1.0) * sigma          # optim/renderer.py:71
    H, W = int(image_size[0]), int(image_size[1])
    return _SoftSilhouette.apply(verts, faces, cam, H, W, float(sigma), float(blur_radius), int(faces_per_pixel), ctx)


class SSRenderer:
    """optim/renderer.py:64-104 with the same constructor arguments and `render` returns: `silhouette_image [1,H,W,4]` (rgb =
    1, alpha = soft silhouette, differentiable w.r.t. the vertices) and `depth [1,H,W,1]` (nearest fragment depth normalised
    to [0,1] over the valid pixels, -1 elsewhere)."""

    def __init__(self, img_shape, h_faces, o_faces, camera_params, device="cuda", ctx: Context | None = None):
        self.img_shape = (int(img_shape[0]), int(img_shape[1]))
        self.h_faces = None if h_faces is None else torch.as_tensor(h_faces).to(device)
        self.o_faces = torch.as_tensor(o_faces).to(device)
        fl = camera_params["focal_length"] if isinstance(camera_params, dict) else camera_params.focal_length
        pp = camera_params["principal_point"] if isinstance(camera_params, dict) else camera_params.principal_point
        self.cam = perspective_camera(torch.as_tensor(fl).detach().cpu(), torch.as_tensor(pp).detach().cpu(), self.img_shape)
        self.sigma = 1e-4
        self.blur_radius = math.log(1.0 / 1e-4 - 1.0) * self.sigma
        self.faces_per_pixel = 100
        self.ctx = ctx

    def render(self, vertices, want_depth=True, **kwargs):
        """want_depth=False skips the depth normalisation (one host synchronisation) and returns depth None."""
        verts, faces = vertices, self.o_faces
        if kwargs.get("h_vertices") is not None:        # join_meshes_as_scene([o_mesh, h_mesh]) (renderer.py:48-61)
            verts = torch.cat([vertices, kwargs["h_vertices"]], 0)
            faces = torch.cat([self.o_faces, self.h_faces + vertices.shape[0]], 0)
        alpha, z = soft_silhouette(verts.float(), faces.to(torch.int32), self.cam, self.img_shape, self.sigma, self.blur_radius,
                                   self.faces_per_pixel, self.ctx)
        image = torch.cat([torch.ones((*alpha.shape, 3), device=alpha.device), alpha[..., None]], -1)[None]
        if not want_depth:
            return image, None
        depth = z.clone()[None, ..., None]
        valid = depth != -1
        if bool(valid.any()):
            d = depth[valid]
            depth[valid] = (d - d.min()) / (d.max() - d.min())
        return image, depth


# ------------------------------------------------------------------------------------------------ pose refinement loop
import torch.nn.functional as _F  # noqa: E402


def matrix_to_rot6d(matrix):
    """optim/utils.py:22-27."""
    matrix = matrix.view(-1, 3, 3)
    return torch.stack((matrix[:, :, 0], matrix[:, :, 1]), dim=-1).view(-1, 6)


def rot6d_to_matrix(rot_6d):
    """optim/utils.py:30-37 (Gram-Schmidt on the two stored columns)."""
    rot_6d = rot_6d.view(-1, 3, 2)
    a1, a2 = rot_6d[:, :, 0], rot_6d[:, :, 1]
    b1 = _F.normalize(a1)
    b2 = _F.normalize(a2 - torch.einsum("bi,bi->b", b1, a2).unsqueeze(-1) * b1)
    b3 = torch.linalg.cross(b1, b2)
    return torch.stack((b1, b2, b3), dim=-1)


def apply_transformation(vertices, rot6d, translation, scaling=1.0):
    """optim/utils.py:56-62: (vertices * scaling) @ R(rot6d) + translation."""
    rot_matrix = rot6d_to_matrix(rot6d).view(1, 3, 3)
    return torch.matmul((vertices * scaling).unsqueeze(1), rot_matrix).squeeze(1) + translation


def calculate_centroid(mask):
    """optim/utils.py:46-53: intensity-weighted (row, col) centroid of the non-zero pixels."""
    # same value without materialising the non-zero coordinates (a data-dependent size = a host synchronisation): zero pixels
    # carry zero weight, and an empty mask falls back to the image centre like the reference
    H, W = mask.shape
    rows = torch.arange(H, device=mask.device, dtype=mask.dtype)
    cols = torch.arange(W, device=mask.device, dtype=mask.dtype)
    tot = mask.sum()
    c = torch.stack([(mask.sum(1) * rows).sum(), (mask.sum(0) * cols).sum()]) / torch.where(tot != 0, tot, torch.ones_like(tot))
    centre = torch.stack([tot.new_full((), H / 2), tot.new_full((), W / 2)])   # fill kernels: no host -> device copy (graph capture)
    return torch.where(tot != 0, c, centre)


def normalized_distance(point1, point2, img_shape):
    """optim/utils.py:40-43."""
    shape = torch.tensor(img_shape, device=point1.device)
    return torch.sqrt(torch.sum((point1 / shape - point2 / shape) ** 2)).item()


class ObjPose_Opt(torch.nn.Module):
    """The optimisation model of optim/optimizer.py:14-175 without its logging side (tensorboard writer, Phong overlay):
    parameters rotation (6-D), translation and optionally scale of the object; `forward(loss_weights)` renders the soft
    silhouette of the transformed object (SSRenderer -> sm_100a kernels) and sums mask (1 - soft "IoU"), centroid and
    contact terms with the reference's weights / kick-in steps.  `human_params` / `object_params` are mappings with
    `vertices`, `contact_verts` (+ `centroid_offset` for the human, `mask` [H,W] for the object)."""

    def __init__(self, rotation_init, translation_init, scaling_init, human_params, object_params, silhouette_renderer,
                 vars=("pose",), ctx: Context | None = None):
        super().__init__()
        g = lambda d, k: d[k] if isinstance(d, dict) else getattr(d, k)
        self.step = 0
        self.ctx = ctx
        self.silhouette_renderer = silhouette_renderer
        self.rotation = torch.nn.Parameter(rotation_init.clone().float(), requires_grad="pose" in vars)
        self.translation = torch.nn.Parameter(translation_init.clone().float(), requires_grad="pose" in vars)
        if "scale" in vars:
            self.scale = torch.nn.Parameter(torch.as_tensor(scaling_init).float(), requires_grad=True)
        else:
            self.register_buffer("scale", torch.as_tensor(scaling_init).float())
        self.register_buffer("human_contact_probs", g(human_params, "contact_verts").float())
        self.register_buffer("object_contact_probs", g(object_params, "contact_verts").float())
        mask = g(object_params, "mask")
        bbox = torch.nonzero(mask)
        lo, hi = bbox.min(dim=0)[0], bbox.max(dim=0)[0]
        self.register_buffer("target_mask_centroid", torch.stack([(hi[0] + lo[0]) / 2, (hi[1] + lo[1]) / 2]).float())
        self.register_buffer("human_vertices", g(human_params, "vertices").float())
        self.register_buffer("hum_centroid_offset", g(human_params, "centroid_offset").float())
        self.register_buffer("obj_vertices", g(object_params, "vertices").float())
        self.register_buffer("target_mask", mask.bool().float())

    def contact_loss(self, obj_verts, human_verts):
        return contact_loss(obj_verts, human_verts, self.object_contact_probs, self.human_contact_probs, self.ctx)

    def mask_loss_iou(self, current_mask):
        """optimizer.py:171-174 (the 'union' is the plain sum of both masks, as in the reference)."""
        return 1 - torch.sum(current_mask * self.target_mask) / torch.sum(current_mask + self.target_mask)

    def forward(self, loss_weights: dict, log: bool = True):
        """log=False leaves out what only the reference's progress bar / tensorboard read (loss floats, centroid distance,
        normalised depth): the iteration then runs without a host synchronisation."""
        obj_vertices = apply_transformation(self.obj_vertices, self.rotation, self.translation, self.scale)
        sil_img, depth_img = self.silhouette_renderer.render(obj_vertices + self.hum_centroid_offset, h_vertices=None, want_depth=log)
        current_mask = sil_img[0, ..., 3]
        active = lambda k: k in loss_weights and self.step >= loss_weights[k]["kick_in"]
        loss_dict = {}
        if active("mask_loss"):
            loss_dict["mask_loss"] = self.mask_loss_iou(current_mask)
        current_mask_centroid = calculate_centroid(current_mask)
        if active("centroid_loss"):
            loss_dict["centroid_loss"] = torch.sum((current_mask_centroid - self.target_mask_centroid) ** 2)
        if active("contact_loss"):
            loss_dict["contact_loss"] = self.contact_loss(obj_vertices, self.human_vertices)
        weighted = {k: v * loss_weights[k]["w"] for k, v in loss_dict.items() if loss_weights[k]["kick_in"] >= 0}
        total = sum(weighted.values())
        self.step += 1
        if not log:
            return total, {"object_vertices": obj_vertices.detach(), "current_mask": current_mask}
        output = {"object_vertices": obj_vertices.detach(), "current_mask": current_mask, "current_depth": depth_img[0, ..., 0],
                  "current_mask_centroid": current_mask_centroid,
                  "centroid_distance": normalized_distance(current_mask_centroid, self.target_mask_centroid, self.target_mask.shape),
                  "losses": {k: float(v.detach()) for k, v in weighted.items()}}
        return total, output


REPLAYED_LAUNCHES = [0]   # kernels of this library launched through CUDA-graph replays (the handles' own counters miss them)


def _adam_groups(model, lr_rotation, lr_translation, lr_scale):
    groups = [{"params": [model.rotation], "lr": lr_rotation}, {"params": [model.translation], "lr": lr_translation}]
    if isinstance(model.scale, torch.nn.Parameter):
        groups.append({"params": [model.scale], "lr": lr_scale})
    return groups


def _fit_graphed(models, loss_weights, max_iter, optimizers):
    """The loop of fit() for one or SEVERAL independent models with each iteration (forward, backward, Adam step: ~150 small
    launches that the host issues in 2-4 ms) captured ONCE per schedule segment into a CUDA graph and replayed.  The set of
    active loss terms only changes at the kick-in steps, so the iterations between two kick-ins are the same launch sequence; the
    first iterations of every segment run eagerly (they are real iterations and warm the allocator / autograd up), one is
    captured, the rest are replays.  Every model has its own side stream; the replays of different models are issued round-robin
    so that their small kernels overlap on the GPU (a 250-iteration fit is a 0.2 s chain of 2-5 us kernels on its own).
    Models advanced together must not share scratch: give each its own ops.Context (the contact term uses the handle's
    workspace)."""
    dev = models[0].rotation.device
    start = models[0].step
    assert all(m.step == start for m in models)
    kicks = {int(w["kick_in"]) - start for w in loss_weights.values() if 0 < int(w.get("kick_in", -1)) - start < max_iter}
    bounds = sorted({0, max_iter} | kicks)
    # eager iterations and the capture share one side stream per model (torch's whole-network capture recipe: the autograd
    # thread's cuBLAS workspaces etc. must exist on the capture stream, or their creation lands on the legacy stream and
    # invalidates the capture)
    cur = torch.cuda.current_stream(dev)
    sides = [torch.cuda.Stream(device=dev) for _ in models]

    def eager(i, k):
        sides[i].wait_stream(cur)
        with torch.cuda.stream(sides[i]):
            for _ in range(k):
                optimizers[i].zero_grad(set_to_none=True)
                loss, _ = models[i](loss_weights, log=False)
                loss.backward()
                optimizers[i].step()

    for a, b in zip(bounds[:-1], bounds[1:]):
        n, warm = b - a, min(3, b - a)
        graphs = []
        for i, model in enumerate(models):
            eager(i, warm)
            if n - warm < 2:
                eager(i, n - warm)
                continue
            step_py = model.step
            g = torch.cuda.CUDAGraph()
            mctx = model.ctx or _ctx(dev)
            l0 = mctx.launch_count()
            optimizers[i].zero_grad(set_to_none=True)
            # capture_begin / capture_end directly: the torch.cuda.graph context manager also runs gc.collect() and
            # torch.cuda.empty_cache(), which hands the model's multi-GB activation arenas back to the driver before every fit
            with torch.cuda.stream(sides[i]):
                g.capture_begin()
                try:
                    loss, _ = model(loss_weights, log=False)
                    loss.backward()
                    optimizers[i].step()
                finally:
                    g.capture_end()
            model.step = step_py             # the capture pass recorded an iteration, it did not run one
            REPLAYED_LAUNCHES[0] += (mctx.launch_count() - l0) * (n - warm - 1)   # launches of this library a replay stands for
            graphs.append((i, g))
        for _ in range(n - warm if graphs else 0):
            for i, g in graphs:
                with torch.cuda.stream(sides[i]):
                    g.replay()
        for i, g in graphs:
            models[i].step += n - warm
        for sd in sides:
            cur.wait_stream(sd)
        del graphs
    assert all(m.step == start + max_iter for m in models)


def fit_many(models, loss_weights: dict, max_iter: int = 250, lr_rotation: float = 5.0e-2, lr_translation: float = 1.0e-2,
             lr_scale: float = 1.0e-2):
    """fit(record=False) for several independent ObjPose_Opt models advanced together (see _fit_graphed): the batch of BASELINE
    configs[3] on one GPU.  Each model needs its own ops.Context."""
    opts = [torch.optim.Adam(_adam_groups(m, lr_rotation, lr_translation, lr_scale), capturable=True) for m in models]
    _fit_graphed(list(models), loss_weights, max_iter, opts)


def fit(model: ObjPose_Opt, loss_weights: dict, max_iter: int = 250, lr_rotation: float = 5.0e-2, lr_translation: float = 1.0e-2,
        lr_scale: float = 1.0e-2, early_stop: bool = False, record: bool = True, graph: bool | None = None):
    """The Adam loop of optim/fit.py:216-290 (per-parameter learning rates of :218-224, optional early stop of :279-283).
    -> list of per-iteration dicts (loss, weighted terms, centroid distance); record=False (no early stop) runs the loop
    without reading anything back and returns an empty list -- by default (graph=None -> True on CUDA) as CUDA-graph replays of
    one captured iteration per schedule segment (_fit_graphed); graph=False issues every iteration from the host."""
    groups = _adam_groups(model, lr_rotation, lr_translation, lr_scale)
    use_graph = (graph if graph is not None else True) and not record and not early_stop and model.rotation.is_cuda
    optimizer = torch.optim.Adam(groups, capturable=True) if use_graph else torch.optim.Adam(groups)
    history, prev = [], 1e10
    if use_graph:
        _fit_graphed([model], loss_weights, max_iter, [optimizer])
        return history
    if not record and not early_stop:
        for _ in range(max_iter):
            optimizer.zero_grad()
            loss, _ = model(loss_weights, log=False)
            loss.backward()
            optimizer.step()
        return history
    for _ in range(max_iter):
        optimizer.zero_grad()
        loss, out = model(loss_weights)
        loss.backward()
        optimizer.step()
        history.append({"loss": loss.item(), "centroid_distance": out["centroid_distance"], **out["losses"]})
        if early_stop and abs(prev - loss.item()) < 1e-6:
            break
        prev = loss.item()
    return history
