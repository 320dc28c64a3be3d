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
