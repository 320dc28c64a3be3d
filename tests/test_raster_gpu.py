"""CUDA rasteriser / Phong shader (ivlm_rasterize_mesh, ivlm_shade_phong through interactvlm_b200.render) against the
CPU oracle (oracle/raster.py) on a real B200.  Integer outputs (pix_to_face, pixel_to_vertices_map) and the barycentrics
are compared BIT-EXACT (the kernel uses explicitly rounded fp32 operations in the oracle's order); the shaded uint8
images may differ by one level where float atomics / fast-math reorder the lighting sums."""
import numpy as np
import pytest
import torch

from interactvlm_b200 import render as R
from interactvlm_b200 import synthetic as S
from interactvlm_b200.ops import LIFT_OBJECT_MESH, LiftMap
from oracle import lift as OL
from oracle import raster as O

pytestmark = pytest.mark.gpu
VIEWS = list(R.OBJECT_VIEWS_4.values())


def _oracle_views(v, f, cams, size):
    out = [O.rasterize(v, f, c, size, size) for c in cams]
    return np.stack([o[0] for o in out]), np.stack([o[1] for o in out]), np.stack([o[2] for o in out]), sum(o[3] for o in out)


@pytest.mark.parametrize("kind,size,views", [("blob", 256, VIEWS), ("torus", 250, VIEWS),
                                             ("adversarial", 256, [(2.0, 0, 0, 0, 0), (1.0, 0, 0, 0, 0), (2.0, 20, 200, 0.1, -0.1)])])
def test_rasteriser_bit_exact_vs_oracle(ctx, kind, size, views):
    v, f = S.make_test_mesh(kind)
    r = R.rasterize_views((v, f), views, (size, size), ctx=ctx)
    from interactvlm_b200 import ops
    z = ops.rasterize_mesh(ctx, r["verts"], r["faces"], r["cams"], size, size, want_p2v=False, want_zbuf=True)
    pix, bary, zb, skipped = _oracle_views(v, f, r["cams"], size)  # same cameras: the comparison is of the rasteriser
    got_pix = r["pix_to_face"].cpu().numpy()
    assert (got_pix >= 0).mean() > 0.05
    assert np.array_equal(got_pix, pix)
    assert np.array_equal(r["bary"].cpu().numpy(), bary)
    assert np.array_equal(z["zbuf"].cpu().numpy(), zb) and np.array_equal(z["pix_to_face"].cpu().numpy(), pix)
    assert r["skipped"] == skipped
    want_p2v = np.where(pix[..., None] >= 0, f[np.maximum(pix, 0)], -1)
    assert np.array_equal(r["p2v"].cpu().numpy(), want_p2v)


def test_rasteriser_non_square_and_pinhole_camera_bit_exact(ctx):
    """Non-square image + the PerspectiveCameras set-up of optim/renderer.py (pixel focal length, off-centre principal
    point): the longer side spans [-W/H, W/H] in NDC; same bit-exact agreement with the oracle."""
    from interactvlm_b200 import ops
    from interactvlm_b200 import optim as PO

    v, f = S.make_test_mesh("torus")
    v = (v * 1.2 + np.array([0.1, -0.05, 2.5])).astype(np.float32)
    H, W = 120, 200
    cam = PO.perspective_camera((150.0, 140.0), (W / 2 + 7.0, H / 2 - 5.0), (H, W))
    vt, ft = torch.from_numpy(v).cuda(), torch.from_numpy(f.astype(np.int32)).cuda()
    r = ops.rasterize_mesh(ctx, vt, ft, [cam], H, W, want_zbuf=True)
    pix, bary, zb, skipped = O.rasterize(v, f, cam, H, W)
    assert (pix >= 0).mean() > 0.05 and np.array_equal(r["pix_to_face"][0].cpu().numpy(), pix)
    assert np.array_equal(r["bary"][0].cpu().numpy(), bary) and np.array_equal(r["zbuf"][0].cpu().numpy(), zb)
    rows, cols = np.nonzero(pix >= 0)
    # OpenCV-style image axes (R = diag(-1,-1,1) undoes pytorch3d's +X-left / +Y-up): world +x and a principal point right of
    # the centre move the object right, world -y and a principal point above the centre move it up
    assert cols.mean() > W / 2 and rows.mean() < H / 2


def test_reference_api_project_vertices_and_create_mask(ctx):
    v, f = S.make_test_mesh("blob")
    v = O.normalize_mesh(v)
    contact = np.nonzero(v[:, 1] > 0.15)[0]
    for p in VIEWS[:2]:
        m, p2v, b = R.project_vertices_and_create_mask((v, f), p, contact, (256, 256), ctx=ctx)
        # the oracle with the product's camera (host trigonometry may differ from numpy's by an ulp)
        pix, bary, _, _ = O.rasterize(v, f, R.camera(p), 256, 256)
        hot = np.isin(f, contact).sum(1) >= 3
        want_m = ((pix >= 0) & hot[np.maximum(pix, 0)]).astype(np.uint8) * 255
        assert m.dtype == np.uint8 and np.array_equal(m, want_m)
        assert p2v.dtype == np.int64 and np.array_equal(p2v, np.where(pix[..., None] >= 0, f[np.maximum(pix, 0)], -1))
        assert b.dtype == np.float32 and np.array_equal(b, bary)


def test_full_size_properties_and_lift_round_trip(ctx):
    """1024^2 x 4 views (the size the reference renders): size-independent properties + Render -> Lift round trip."""
    v, f = S.make_test_mesh("blob", n_lat=96, n_lon=192)          # 36 k faces
    v = O.normalize_mesh(v)
    r = R.rasterize_views((v, f), VIEWS, (1024, 1024), ctx=ctx)
    pix, bary, p2v = r["pix_to_face"], r["bary"], r["p2v"]
    fg = pix >= 0
    assert 0.05 < fg.float().mean().item() < 0.6 and r["skipped"] == 0
    assert (bary[fg].sum(-1) - 1).abs().max().item() < 1e-5 and bary[fg].min().item() > 0
    assert bool((bary[~fg] == -1).all()) and bool((p2v[~fg] == -1).all())
    ft = torch.from_numpy(f).to(pix.device)
    assert torch.equal(p2v[fg], ft[pix[fg].long()])
    # every pixel's interpolated surface point projects back onto the pixel centre
    P = (bary[fg][:, :, None] * r["verts"][p2v[fg]]).sum(1).cpu().numpy()
    vi, rows, cols = [t.cpu().numpy() for t in torch.nonzero(fg, as_tuple=True)]
    for k in range(4):
        sel = vi == k
        ndc = O.project(P[sel].astype(np.float32), r["cams"][k])
        assert np.abs(ndc[:, 0] - (1 - (2 * cols[sel] + 1) / 1024)).max() < 2e-5
        assert np.abs(ndc[:, 1] - (1 - (2 * rows[sel] + 1) / 1024)).max() < 2e-5
    # contact region -> masks -> lift
    contact = np.nonzero(v[:, 1] > 0.15)[0]
    hot = torch.from_numpy(np.isin(f, contact).all(1)).to(pix.device)
    mask = fg & hot[pix.clamp_min(0).long()]
    logits = torch.where(mask, 12.0, -12.0).float()[None].contiguous()
    lm = R.lift_map_from_mesh((v, f), VIEWS, (1024, 1024), ctx=ctx)
    got = lm(logits, LIFT_OBJECT_MESH).cpu().numpy()[0]
    want = OL.lift_object_mesh(logits.cpu().numpy(), p2v.cpu().numpy(), bary.cpu().numpy(), len(v))[0]
    # vertices of this coarse mesh collect thousands of pixels: the oracle's sequential fp32 sum (the reference's scatter order)
    # and the kernel's lane-strided + shuffle-tree sum differ by a few 1e-7 per thousand terms
    assert np.abs(got - want).max() < 5e-6
    visible = np.zeros(len(v), bool)
    visible[np.unique(p2v[fg].cpu().numpy())] = True
    interior = np.zeros(len(v), bool)
    interior[contact] = True
    interior[np.unique(f[~np.isin(f, contact).all(1)])] = False
    sel = interior & visible
    assert sel.sum() > 1000 and (got[sel] > 0.5).all() and (got[~np.isin(np.arange(len(v)), contact)] < 0.5).all()


def test_phong_render_vs_oracle(ctx):
    v, f = S.make_test_mesh("blob", n_lat=40, n_lon=80)
    v = O.normalize_mesh(v)
    col = ((v - v.min(0)) / (v.max(0) - v.min(0)) * 0.8 + 0.1).astype(np.float32)
    for p, light in zip(VIEWS[:2], R.LIGHT_LOCATIONS[:2]):
        img = R.render_mesh((v, f), p, light, (256, 256), vertex_colors=col, ctx=ctx)
        want = O.render_phong(v, f, col, R.camera(p), light, 256, 256)
        assert img.dtype == np.uint8 and img.shape == (256, 256, 3)
        d = np.abs(img.astype(int) - want.astype(int))
        assert (img[0, 0] == 255).all() and d.max() <= 2 and (d > 0).mean() < 0.02, (d.max(), (d > 0).mean())
        assert 20 < img[want.sum(-1) < 765].mean() < 250


def test_generate_sam_inp_objs_and_lift_from_pickle(ctx, tmp_path):
    """utils/demo_utils.py:171-256 flow: obj file -> 8 PNGs + lift2d_dict.pkl that the object predictor consumes."""
    import joblib

    v, f = S.make_test_mesh("torus")
    obj = tmp_path / "thing" / "mesh.obj"
    obj.parent.mkdir()
    with open(obj, "w") as fh:
        fh.writelines(f"v {x[0]:.7f} {x[1]:.7f} {x[2]:.7f}\n" for x in v * 2.5 + 1.0)
        fh.writelines(f"f {t[0] + 1} {t[1] + 1} {t[2] + 1}\n" for t in f)
    out = R.generate_sam_inp_objs(str(obj), image_size=(256, 256))
    names = sorted(p.name for p in out.iterdir())
    assert len(names) == 9 and "lift2d_dict.pkl" in names and "obj_render_grey_frontleft.png" in names
    d = joblib.load(out / "lift2d_dict.pkl")
    assert d["num_vertices"] == len(v) and len(d["pixel_to_vertices_map"]) == 4
    p2v, bary = np.stack(d["pixel_to_vertices_map"]), np.stack(d["bary_coords_map"])
    assert p2v.dtype == np.int64 and p2v.shape == (4, 256, 256, 3) and bary.dtype == np.float32
    logits = torch.from_numpy(S.make_mask_logits(1, size=256, seed=3)).cuda()
    got = LiftMap(ctx, p2v, bary, len(v))(logits, LIFT_OBJECT_MESH).cpu().numpy()
    want = OL.lift_object_mesh(logits.cpu().numpy(), p2v, bary, len(v))
    assert np.abs(got - want).max() < 1e-6


def test_object_demo_flow_mesh_to_vertex_contact(ctx, tmp_path):
    """run_demo.py's object path end to end on the tiny configuration: .obj -> rendered views + lift2d_dict.pkl (GPU
    rasteriser) -> evaluate(contact_type='oafford') -> per-vertex contact == oracle lift of the predicted masks."""
    import importlib.util
    import joblib

    from interactvlm_b200.config import IVLMConfig
    from interactvlm_b200.model import InteractVLMForCausalLM

    spec = importlib.util.spec_from_file_location("demo_obj", str(R.Path(__file__).resolve().parents[1] / "examples" / "demo_synthetic_object.py"))
    demo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(demo)
    cfg = IVLMConfig.tiny()
    model = InteractVLMForCausalLM(cfg, S.make_state_dict(cfg, seed=0), ctx=ctx)
    _, res, f, obj_dir = demo.run(tmp_path, "tiny", model=model)
    d = joblib.load(obj_dir / "lift2d_dict.pkl")
    p2v, bary = np.stack(d["pixel_to_vertices_map"]), np.stack(d["bary_coords_map"])
    masks = res["pred_masks"][0].cpu().numpy()[None]
    assert masks.shape == (1, 4, 1024, 1024)
    want = OL.lift_object_mesh(masks, p2v, bary, d["num_vertices"])
    got = np.load(f)["pred_contact_3d"]
    assert got.shape == (1, d["num_vertices"]) and np.abs(got - want).max() < 5e-6   # tree vs sequential fp32 sums
    assert np.array_equal(got > 0.5, want > 0.5)


def test_point_cloud_rasteriser_vs_oracle(ctx):
    """ivlm_rasterize_points (the p2pmap producer of preprocess_data/utils_obj_pc.py:88-113) bit-exact against the numpy
    oracle, plus the Render -> Lift round trip for a 2048-point cloud: per-point affordance -> masks -> lift recovers it."""
    from interactvlm_b200 import render as R
    from interactvlm_b200.ops import LIFT_POINTS, LiftMap
    from oracle import raster as O

    g = np.random.default_rng(0)
    d = g.normal(size=(2048, 3))
    pts = (d / np.linalg.norm(d, axis=1, keepdims=True) * (0.35 + 0.05 * np.sin(5 * d[:, :1]))).astype(np.float32)
    views = [(2.0, 45.0, 315.0, 0.0, 0.0), (2.0, 315.0, 135.0, 0.0, 0.3)]
    maps = []
    for cp in views:
        got = R.project_points_to_image(pts, cp, dynamic_radius=True, image_size=(256, 256), ctx=ctx)
        want = O.rasterize_points(pts, O.camera((cp[0], cp[1], cp[2], 0.0, cp[4])), 256, 256, R.get_dynamic_radius(pts))
        assert got.dtype == np.int64 and got.shape == (256, 256)
        assert np.array_equal(got, want)
        assert 0.02 < (got >= 0).mean() < 0.6 and got.max() < 2048
        maps.append(got)
    # round trip: affordance of the visible points -> heat masks -> lift -> the same values on every point seen
    afford = g.random(2048).astype(np.float32)
    p2p = np.stack(maps)
    masks = np.where(p2p >= 0, afford[np.maximum(p2p, 0)], 0.0).astype(np.float32)[None]
    lm = LiftMap(ctx, p2p, None, 2048)
    out = lm(torch.from_numpy(masks).cuda(), LIFT_POINTS).cpu().numpy()[0]
    seen = np.zeros(2048, bool)
    seen[np.unique(p2p[p2p >= 0])] = True
    assert np.abs(out[seen] - afford[seen]).max() < 1e-6 and np.all(out[~seen] == 0)
    mask, _ = R.create_affordance_mask(pts, np.nonzero(afford > 0.5)[0], views[0], image_size=(256, 256), ctx=ctx)
    assert np.array_equal(mask == 255, (maps[0] >= 0) & (afford[np.maximum(maps[0], 0)] > 0.5))
