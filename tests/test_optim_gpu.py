"""Fused contact-loss kernel (ivlm_contact_loss through interactvlm_b200.optim) on a real B200 against the goldens recorded
from the reference's own function and against the float64 oracle.  Tolerances: the kernel is fp32 with direct differences;
the reference's own fp32 run deviates 2e-6 from its fp64 run on these inputs -- ours must be at least that close."""
from pathlib import Path

import numpy as np
import pytest
import torch

from interactvlm_b200 import optim as PO
from oracle import optim as OO
from oracle.make_goldens_optim import CASES, inputs

pytestmark = pytest.mark.gpu
GOLD = np.load(Path(__file__).parent / "golden" / "contact_loss.npz")


def case(name):
    seed, n_obj, n_hum = CASES[name]
    obj, hum, p, q = inputs(seed, n_obj, n_hum)
    if name == "coincident":
        hum[:16] = obj[:16]
    return obj, hum, p, q


@pytest.mark.parametrize("name", list(CASES))
def test_value_and_gradient_vs_reference_goldens(ctx, name):
    obj, hum, p, q = case(name)
    o = torch.from_numpy(obj).cuda().requires_grad_(True)
    n0 = ctx.launch_count()
    loss = PO.contact_loss(o, torch.from_numpy(hum).cuda(), torch.from_numpy(p).cuda(), torch.from_numpy(q).cuda(), ctx=ctx)
    assert ctx.launch_count() - n0 == 2
    (3.0 * loss).backward()
    g64 = GOLD[f"{name}_f64_grad"]
    assert abs(loss.item() - GOLD[f"{name}_f64_loss"]) < 2e-6 * max(1.0, abs(GOLD[f"{name}_f64_loss"]))
    assert np.abs(o.grad.cpu().numpy() / 3.0 - g64).max() < 2e-6 * np.abs(g64).max() + 1e-9
    assert abs(loss.item() - GOLD[f"{name}_f32_loss"]) < 4e-6   # and next to the reference's own fp32 result


def test_inside_the_reference_transformation_chain(ctx):
    """Gradient flows through apply_transformation-style ops (optim/utils.py:56-62) to rotation / translation / scale."""
    obj, hum, p, q = case("small")
    dev = "cuda"
    R6 = torch.tensor([[1.0, 0.1, 0.0], [0.0, 1.0, 0.2], [0.1, 0.0, 1.0]], device=dev, requires_grad=True)
    t = torch.tensor([0.1, -0.2, 0.3], device=dev, requires_grad=True)
    s = torch.tensor(1.3, device=dev, requires_grad=True)
    o0, h_, p_, q_ = (torch.from_numpy(a).to(dev) for a in (obj, hum, p, q))
    loss = PO.contact_loss((o0 * s) @ R6 + t, h_, p_, q_, ctx=ctx)
    loss.backward()
    R6d, td, sd = (x.detach().double().cpu().requires_grad_(True) for x in (R6, t, s))
    od = (torch.from_numpy(obj).double() * sd) @ R6d + td
    dist = torch.cdist(od[None], torch.from_numpy(hum).double()[None])[0]
    w = torch.outer(torch.from_numpy(p).double(), torch.from_numpy(q).double())
    ref = (dist * w).sum() / w.sum()
    ref.backward()
    assert abs(loss.item() - ref.item()) < 2e-6
    for a, b in ((R6, R6d), (t, td), (s, sd)):
        assert (a.grad.double().cpu() - b.grad).abs().max().item() < 1e-5 * max(1.0, b.grad.abs().max().item())


def test_full_size_properties(ctx):
    """20 k object vertices x SMPL-X (10475): the reference would build two 0.8 GB matrices; size-independent properties."""
    g = torch.Generator().manual_seed(0)
    o = (torch.randn(20000, 3, generator=g) * 0.3).cuda()
    h = (torch.randn(10475, 3, generator=g) * 0.4).cuda()
    p, q = torch.rand(20000, generator=g).cuda(), torch.rand(10475, generator=g).cuda()
    base = PO.contact_loss(o, h, p, q, ctx=ctx).item()
    shift = torch.tensor([0.3, -1.0, 2.0], device="cuda")
    assert abs(PO.contact_loss(o + shift, h + shift, p, q, ctx=ctx).item() - base) < 1e-5          # translation invariance
    assert abs(PO.contact_loss(o * 2.5, h * 2.5, p, q, ctx=ctx).item() - 2.5 * base) < 1e-5 * 2.5  # homogeneity
    assert abs(PO.contact_loss(o, h, p * 3, q * 0.5, ctx=ctx).item() - base) < 1e-6                # weights are normalised
    assert PO.contact_loss(o, h, p, q, ctx=ctx).item() == base                                      # deterministic
    sub = slice(0, 2000)
    want, gw = OO.contact_loss(o[sub].cpu().numpy(), h.cpu().numpy(), p[sub].cpu().numpy(), q.cpu().numpy())
    os_ = o[sub].clone().requires_grad_(True)
    got = PO.contact_loss(os_, h, p[sub], q, ctx=ctx)
    got.backward()
    assert abs(got.item() - want) < 2e-6 and np.abs(os_.grad.cpu().numpy() - gw).max() < 2e-6 * np.abs(gw).max() + 1e-10


def test_knn1_exact_vs_brute_force(ctx):
    g = torch.Generator().manual_seed(5)
    for n, m, D in ((1000, 777, 6), (513, 4000, 3), (7, 1, 8), (300, 300, 1)):
        x, y = torch.randn(n, D, generator=g).cuda(), torch.randn(m, D, generator=g).cuda()
        y[m // 2] = y[0]                                     # an exact duplicate: ties go to the lower index
        idx, d2 = PO.knn1(x, y, ctx)
        full = ((x.double()[:, None, :] - y.double()[None, :, :]) ** 2).sum(-1)
        want_d, want_i = full.min(1)
        assert (d2.double() - want_d).abs().max().item() < 1e-5
        picked = full.gather(1, idx[:, None])[:, 0]
        assert (picked - want_d).abs().max().item() < 1e-5   # the chosen neighbour is a nearest one (fp32 near-ties aside)
        assert (idx != m // 2).all() or m == 1


def test_icp_vs_reference_goldens(ctx):
    from oracle.make_goldens_optim import ICP_CASES, icp_inputs

    G = np.load(Path(__file__).parent / "golden" / "icp.npz")
    for name, (seed, n_obj, n_hum, est) in ICP_CASES.items():
        obj, on, hum, hn, R0, T0, s0 = icp_inputs(seed, n_obj, n_hum)
        t = lambda a: torch.from_numpy(a).cuda()
        sol = PO.ICP(t(obj), t(hum), init_transform=PO.SimilarityTransform(t(R0), t(T0), t(s0)), max_iterations=30,
                     estimate_scale=est, obj_contact_normals=t(on), hum_contact_normals=t(hn), ctx=ctx)
        assert sol.converged == bool(G[f"{name}_f64_converged"]) and len(sol.t_history) == int(G[f"{name}_f64_iters"])
        assert np.abs(sol.RTs.R.cpu().numpy() - G[f"{name}_f64_R"]).max() < 2e-5
        assert np.abs(sol.RTs.T.cpu().numpy() - G[f"{name}_f64_T"]).max() < 2e-5
        assert np.abs(sol.RTs.s.cpu().numpy() - G[f"{name}_f64_s"]).max() < 2e-5
        assert np.abs(sol.Xt.cpu().numpy() - G[f"{name}_f64_Xt"]).max() < 2e-5
        assert np.abs(sol.rmse.cpu().numpy() - G[f"{name}_f64_rmse"]).max() < 2e-5
        # the reference's own fp32 run is this far from its fp64 run:
        assert np.abs(G[f"{name}_f32_Xt"] - G[f"{name}_f64_Xt"]).max() < 2e-5


def _sil_case(H=48, W=64):
    from interactvlm_b200 import synthetic as S
    from oracle import raster as OR

    v, f = S.make_test_mesh("blob", n_lat=12, n_lon=20)
    v = (v * 0.7 + np.array([0.08, -0.05, 2.2])).astype(np.float32)
    fl, pp = (80.0, 78.0), (W / 2 + 3.0, H / 2 - 2.0)
    return v, f, fl, pp, OR.perspective_camera(fl, pp, (H, W)), (H, W)


def test_soft_silhouette_forward_backward_vs_oracle(ctx):
    """ivlm_soft_silhouette(+_backward) against the float64 oracle (whose gradient is checked numerically on the CPU side)."""
    from oracle import raster as OR

    v, f, fl, pp, ocam, (H, W) = _sil_case()
    cam = PO.perspective_camera(fl, pp, (H, W))
    for k in ("fx", "fy", "cx", "cy"):
        assert abs(float(cam[k]) - float(ocam[k])) < 1e-6
    G = np.random.default_rng(1).normal(size=(H, W)).astype(np.float32)
    want_a, want_z, want_g = OR.soft_silhouette(v.astype(np.float64), f, ocam, H, W, grad_alpha=G.astype(np.float64))
    vt = torch.from_numpy(v).cuda().requires_grad_(True)
    ft = torch.from_numpy(f.astype(np.int32)).cuda()
    alpha, z = PO.soft_silhouette(vt, ft, cam, (H, W), ctx=ctx)
    (alpha * torch.from_numpy(G).cuda()).sum().backward()
    a = alpha.detach().cpu().numpy()
    assert np.abs(a - want_a).max() < 2e-3 and np.abs(a - want_a).mean() < 2e-5      # fp32 distances / sigma = 1e-4
    zz = z.cpu().numpy()
    both = (zz >= 0) & (want_z >= 0)
    assert (both == (want_z >= 0)).mean() > 0.999 and np.abs(zz[both] - want_z[both]).max() < 1e-4
    g = vt.grad.cpu().numpy()
    assert np.abs(g - want_g).max() < 2e-2 * np.abs(want_g).max()
    assert np.abs(g - want_g).sum() < 5e-3 * np.abs(want_g).sum()
    # fewer fragments than candidates: the K nearest in depth are kept
    want_k, _ = OR.soft_silhouette(v.astype(np.float64), f, ocam, H, W, K=3)
    got_k, _ = PO.soft_silhouette(vt.detach(), ft, cam, (H, W), faces_per_pixel=3, ctx=ctx)
    dk = np.abs(got_k.cpu().numpy() - want_k)
    # depth ties (two faces clipped onto their shared edge give the same pz) are broken by rounding, so a few boundary pixels
    # may keep a different third fragment in fp32 than in fp64; everywhere else the truncated lists agree
    assert (dk > 2e-3).mean() < 0.02 and np.median(dk) < 1e-6 and np.abs(want_k - want_a).max() > 1e-3, ((dk > 2e-3).mean(), dk.max())


def test_ssrenderer_mask_loss_drives_translation(ctx):
    """SSRenderer.render + mask_loss_iou (optim/optimizer.py:171-174) inside torch autograd: a few Adam steps on a
    translation move the silhouette onto the target mask."""
    v, f, fl, pp, _, (H, W) = _sil_case()
    ren = PO.SSRenderer((H, W), None, torch.from_numpy(f), {"focal_length": torch.tensor(fl), "principal_point": torch.tensor(pp)}, ctx=ctx)
    base = torch.from_numpy(v).cuda()
    with torch.no_grad():
        target_img, depth = ren.render(base)
        target = (target_img[0, ..., 3] > 0.5).float()
    assert target_img.shape == (1, H, W, 4) and depth.shape == (1, H, W, 1) and float(depth.max()) == 1.0 and float(depth.min()) == -1.0
    t = torch.tensor([0.12, -0.08, 0.0], device="cuda", requires_grad=True)
    opt = torch.optim.Adam([t], lr=0.01)

    def iou_loss():
        cur = ren.render(base + t)[0][0, ..., 3]
        return 1 - (cur * target).sum() / (cur + target).sum()      # the reference's "IoU" (union = sum of both)

    first = iou_loss().item()
    for _ in range(60):
        opt.zero_grad()
        loss = iou_loss()
        loss.backward()
        opt.step()
    assert loss.item() < first - 0.05 and t.detach().abs().max().item() < 0.06, (first, loss.item(), t)


def test_pose_refinement_end_to_end(ctx):
    """optim/fit.py's flow on a synthetic scene (examples/demo_synthetic_fit.py): contact ICP initialisation, then Adam on
    rotation / translation with the mask + centroid + contact terms; the loss falls and the object moves towards the pose
    that produced the target mask."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("demo_fit", str(Path(__file__).resolve().parents[1] / "examples" / "demo_synthetic_fit.py"))
    demo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(demo)
    model, hist, info = demo.run(iters=100, size=96)
    assert len(hist) == 100 and all(np.isfinite(h["loss"]) for h in hist)
    # the reference's mask term is 1 - I / (A + B) >= 0.5 and the contact term is a mean distance: both have a floor
    assert hist[-1]["loss"] < 0.8 * hist[0]["loss"], (hist[0], hist[-1])
    assert hist[-1]["mask_loss"] < hist[0]["mask_loss"] and hist[-1]["contact_loss"] < hist[0]["contact_loss"]
    assert hist[-1]["centroid_distance"] < 0.3 * hist[0]["centroid_distance"]
    assert info["t_err_end"] < 0.25 * info["t_err_start"] and info["t_err_end"] < 0.04, info


def test_fit_driver_on_a_synthetic_scene(ctx):
    """interactvlm_b200.fit.run_fit (optim/fit.py:86-290): contact thresholds -> mask-centroid translation -> normal filter ->
    contact ICP -> Adam loop.  The object must end closer to its true pose than the ICP initialisation, and the loop without
    host read-backs (record=False) must produce the same parameters as the recording one."""
    from interactvlm_b200 import fit as FIT
    from interactvlm_b200.bench_fit import _scene

    human, obj, cam = _scene(128, 3, torch.device("cuda"))
    gt_t = torch.tensor([0.62, 0.1, 2.75], device="cuda")
    opt = FIT.default_options()
    opt["max_iter"] = 80
    opt["init"]["translation_hum_centroid"] = False      # the reference's mask-pixel pick is far off for this camera; start from ICP
    res = FIT.run_fit(human, obj, cam, (128, 128), opt, record=True)
    assert len(res.history) == 80 and res.history[-1]["loss"] < res.history[0]["loss"]
    err_end = float((res.translation - gt_t).norm())
    assert err_end < 0.25, err_end
    # record=False replays ONE captured iteration as a CUDA graph (optim._fit_graphed).  The soft-silhouette backward accumulates
    # per-vertex gradients with float atomics and the loop amplifies the differences (two EAGER runs differ by ~7e-4 after 80
    # steps, 1e-5 after 40): the loops are compared after 30 steps, the converged poses loosely
    res2 = FIT.run_fit(human, obj, cam, (128, 128), opt, record=False)
    assert res2.history == [] and torch.allclose(res2.translation, res.translation, atol=8e-3)
    assert torch.allclose(res2.rotation6d, res.rotation6d, atol=8e-3)
    assert float((res2.translation - gt_t).norm()) < 0.25
    opt["max_iter"] = 30
    a = FIT.run_fit(human, obj, cam, (128, 128), opt, record=True)
    b = FIT.run_fit(human, obj, cam, (128, 128), opt, record=False)
    assert torch.allclose(a.translation, b.translation, atol=3e-4) and torch.allclose(a.rotation6d, b.rotation6d, atol=3e-4)


def test_fits_advanced_together_equal_separate_fits(ctx):
    """fit.run_fit_many (one handle + stream + captured iteration per scene, interleaved graph replays) against run_fit per scene."""
    from interactvlm_b200 import fit as FIT
    from interactvlm_b200.bench_fit import _scene

    dev = torch.device("cuda")
    scenes = [_scene(128, 3 + i, dev) for i in range(3)]
    opt = FIT.default_options()
    opt["max_iter"] = 30
    opt["init"]["translation_hum_centroid"] = False
    many = FIT.run_fit_many(scenes, (128, 128), opt)
    for (h, o, c), r in zip(scenes, many):
        one = FIT.run_fit(h, o, c, (128, 128), opt, record=False)
        assert torch.allclose(one.translation, r.translation, atol=3e-4) and torch.allclose(one.rotation6d, r.rotation6d, atol=3e-4)
    assert len({tuple(round(float(x), 3) for x in r.translation) for r in many}) == 3      # three different scenes, three poses
