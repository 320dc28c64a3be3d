"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the contact term of the reference's pose refinement,
ObjPose_Opt.contact_loss (optim/optimizer.py:80-96): weighted mean of the pairwise object-human vertex distances with the
outer product of the two contact-probability vectors as weights.  Pinned: tests/golden/contact_loss.npz holds outputs of
the reference's own function body (extracted from /root/reference/optim/optimizer.py by oracle/make_goldens_optim.py and
executed unmodified, value and autograd gradient); tests/test_optim_cpu.py checks this restatement against them.
Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module."""
from __future__ import annotations

import numpy as np


def contact_loss(obj_verts, human_verts, obj_probs, human_probs, dtype=np.float64):
    """-> (loss, d loss / d obj_verts).  optimizer.py:82-96: dist = cdist(obj, human); w = outer(p_obj, p_human);
    loss = (dist * w).sum() / w.sum().  Direct differences (no |a|^2 + |b|^2 - 2ab expansion), `dtype` arithmetic."""
    o, h = np.asarray(obj_verts, dtype=dtype), np.asarray(human_verts, dtype=dtype)
    p, q = np.asarray(obj_probs, dtype=dtype), np.asarray(human_probs, dtype=dtype)
    loss_num = 0.0
    grad = np.zeros_like(o)
    for i0 in range(0, len(o), 1024):  # row blocks keep the pair matrix small
        diff = o[i0:i0 + 1024, None, :] - h[None, :, :]
        dist = np.sqrt((diff * diff).sum(-1))
        loss_num += float((dist * q[None, :]).sum(1) @ p[i0:i0 + 1024])
        inv = np.divide(1.0, dist, out=np.zeros_like(dist), where=dist > 0)
        grad[i0:i0 + 1024] = ((q[None, :] * inv)[..., None] * diff).sum(1) * p[i0:i0 + 1024, None]
    denom = p.sum() * q.sum()
    return loss_num / denom, grad / denom
