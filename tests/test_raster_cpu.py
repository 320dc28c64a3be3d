"""CPU checks of the "Render" step: the oracle (oracle/raster.py, pytorch3d's published algorithm -- PARITY UNPINNED, the
library is absent) against analytic cases and the conventions the reference relies on, and the host-side camera /
mesh helpers of interactvlm_b200.render against the oracle."""
import numpy as np
import pytest
import torch

from interactvlm_b200 import render as R
from interactvlm_b200 import synthetic as S
from oracle import raster as O

VIEWS = list(R.OBJECT_VIEWS_4.values())


def test_look_at_is_orthonormal_and_looks_at_origin():
    for p in VIEWS + [(2.0, 0, 0, 0, 0), (2.0, 90, 0, 0, 0), (3.0, -30, 200, 0, 0)]:
        c = O.camera(p)
        assert np.allclose(c["R"].T @ c["R"], np.eye(3), atol=1e-6)
        assert abs(np.linalg.norm(c["C"]) - p[0]) < 1e-5
        origin = O.project(np.zeros((1, 3), np.float32), c)[0]
        assert abs(origin[0]) < 1e-6 and abs(origin[1]) < 1e-6 and abs(origin[2] - p[0]) < 1e-5  # centre of the image, depth = dist


def test_host_camera_matches_oracle():
    for p in VIEWS + [(2.0, 10, 20, 0.1, -0.2)]:
        a, b = R.camera(p), O.camera(p)
        for k in ("R", "T", "C"):
            assert np.abs(np.asarray(a[k]) - b[k]).max() < 1e-6
        assert abs(float(a["s"]) - float(b["s"])) < 1e-6 and float(a["z_clip"]) == float(b["z_clip"]) == 0.5


def test_fronto_parallel_triangle_is_analytic():
    v = np.array([[-0.5, -0.5, 0], [0.5, -0.5, 0], [0, 0.5, 0]], dtype=np.float32)
    f = np.array([[0, 1, 2]])
    H = 128
    pix, bary, z, skipped = O.rasterize(v, f, O.camera((2.0, 0, 0, 0, 0)), H, H)
    assert skipped == 0 and np.allclose(z[pix >= 0], 2.0, atol=1e-5)
    # pinhole area: triangle of base 1, height 1 at depth 2 with fov 60: ndc = x * sqrt(3) / 2 -> pixels = ndc * H / 2
    expect = 0.5 * (np.sqrt(3) / 2 * H / 2) ** 2
    assert abs((pix >= 0).sum() - expect) < 0.03 * expect
    fg = bary[pix >= 0]
    assert np.abs(fg.sum(-1) - 1).max() < 1e-6 and fg.min() > 0
    assert (bary[pix < 0] == -1).all()                       # pytorch3d fills the background with -1
    rows, cols = np.nonzero(pix >= 0)
    assert rows[np.argmax(fg[:, 2])] == rows.min()           # +Y is up: the apex (vertex 2) is in the top rows
    assert cols[np.argmax(fg[:, 1])] > cols[np.argmax(fg[:, 0])]  # from the front, world +X is on the right
    # interpolated world position at every pixel reproduces the pixel ray (perspective-correct barycentrics)
    P = fg @ v
    c = O.camera((2.0, 0, 0, 0, 0))
    ndc = O.project(P.astype(np.float32), c)
    ys = 1 - (2 * rows + 1) / H
    xs = 1 - (2 * cols + 1) / H
    assert np.abs(ndc[:, 0] - xs).max() < 1e-5 and np.abs(ndc[:, 1] - ys).max() < 1e-5


def test_depth_order_ties_and_clipping():
    v, f = S.make_test_mesh("adversarial")
    nf = len(f)
    pix, bary, z, skipped = O.rasterize(v, f, O.camera((2.0, 0, 0, 0, 0)), 256, 256)
    seen = lambda i: int((pix == i).sum())
    assert seen(nf - 10) > 0 and seen(nf - 9) > 0            # quad A in front of the blob
    assert seen(nf - 2) == 0 and seen(nf - 1) == 0           # exact duplicates lose the tie to the lower face index
    assert seen(nf - 6) == 0 and seen(nf - 5) == 0           # degenerate face, face behind the camera
    assert seen(nf - 4) > 0 and z[pix == nf - 4].min() >= 0.5  # straddling face kept only beyond z_clip
    assert seen(nf - 3) > 0                                   # partially outside the view
    assert skipped == 0
    # a closer camera: the large face now crosses the camera plane itself -> skipped and counted
    _, _, _, skipped = O.rasterize(v, f, O.camera((1.0, 0, 0, 0, 0)), 64, 64)
    assert skipped > 0


def test_project_vertices_and_create_mask_round_trip():
    """GT contact vertices -> per-view masks -> lift (the reference's verify_contact_reconstruction_diff idea,
    render_mesh_utils.py:200-238): every contact vertex that is visible comes back."""
    from oracle import lift as OL

    v, f = S.make_test_mesh("blob")
    v = O.normalize_mesh(v)
    contact = np.nonzero(v[:, 1] > 0.15)[0]
    size = 256
    masks, p2vs, barys = [], [], []
    for p in VIEWS:
        m, p2v, b = O.project_vertices_and_create_mask(v, f, p, contact, (size, size))
        assert m.dtype == np.uint8 and set(np.unique(m)) <= {0, 255}
        assert p2v.dtype == np.int64 and p2v.shape == (size, size, 3) and b.shape == (size, size, 3)
        assert ((p2v >= 0).all(-1) == (p2v >= 0).any(-1)).all()
        masks.append(m), p2vs.append(p2v), barys.append(b)
    logits = np.where(np.stack(masks) > 0, 12.0, -12.0).astype(np.float32)[None]
    got = OL.lift_object_mesh(logits, np.stack(p2vs), np.stack(barys), len(v))[0]
    visible = np.zeros(len(v), bool)
    for p2v in p2vs:
        visible[p2v[p2v >= 0]] = True
    interior = np.zeros(len(v), bool)  # contact vertices all of whose faces are contact faces
    hot = np.isin(f, contact).all(1)
    interior[contact] = True
    for face, h in zip(f, hot):
        if not h:
            interior[face] = False
    sel = interior & visible
    assert sel.sum() > 50 and (got[sel] > 0.5).all()
    assert (got[~np.isin(np.arange(len(v)), contact)] < 0.5).all()


def test_normalize_mesh_and_obj_reader(tmp_path):
    v, f = S.make_test_mesh("torus")
    v = v * 3.7 + np.array([1.0, -2.0, 0.5], np.float32)
    a, b = R.normalize_mesh(torch.from_numpy(v)).numpy(), O.normalize_mesh(v)
    assert np.abs(a - b).max() < 1e-6 and abs((b.max(0) - b.min(0)).max() - 1) < 1e-6 and np.abs(b.mean(0)).max() < 1e-6
    path = tmp_path / "m.obj"
    with open(path, "w") as fh:
        fh.write("# test\n")
        for x in v:
            fh.write(f"v {x[0]:.8f} {x[1]:.8f} {x[2]:.8f}\n")
        fh.write("vt 0 0\nvn 0 0 1\n")
        for t in f[:-2]:
            fh.write(f"f {t[0] + 1}/1/1 {t[1] + 1}/1/1 {t[2] + 1}/1/1\n")
        q = f[-2:]
        fh.write(f"f {q[0][0] + 1} {q[0][1] + 1} {q[0][2] + 1} {q[1][2] + 1}\n")  # a quad: fan-triangulated
        fh.write(f"f -3 -2 -1\n")
    lv, lf = R.load_obj(path)
    assert np.allclose(lv.numpy(), v, atol=1e-6) and lf.shape == (len(f) + 1, 3)
    assert np.array_equal(lf[:-3].numpy(), f[:-2]) and lf[-1].tolist() == [len(v) - 3, len(v) - 2, len(v) - 1]


def test_render_module_has_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        R.rasterize_views(S.make_test_mesh("blob"), VIEWS, (64, 64))


def test_soft_silhouette_oracle_gradient_is_the_numerical_one():
    """The oracle's analytic backward of the soft silhouette (pytorch3d's formulas) against central finite differences of its
    own forward, on a non-square image with an off-centre principal point."""
    v, f = S.make_test_mesh("blob", n_lat=8, n_lon=12)
    v = (v * 0.6 + np.array([0.1, -0.05, 2.0])).astype(np.float64)
    cam = O.perspective_camera((60.0, 60.0), (24.0, 16.0), (32, 48))
    G = np.random.default_rng(0).normal(size=(32, 48))
    alpha, z, g = O.soft_silhouette(v, f, cam, 32, 48, grad_alpha=G)
    assert alpha.min() >= 0 and alpha.max() <= 1 and 0.05 < (alpha > 0.5).mean() < 0.5
    assert ((alpha > 0.01) & (alpha < 0.99)).sum() > 30              # a soft boundary band exists
    assert (z[alpha > 0.5] > 1.5).all() and (z[alpha == 0] == -1).all()
    loss = lambda vv: (O.soft_silhouette(vv, f, cam, 32, 48)[0] * G).sum()
    for k in np.argsort(-np.abs(g).reshape(-1))[:5]:
        i, d = divmod(int(k), 3)
        vp, vm = v.copy(), v.copy()
        vp[i, d] += 1e-6
        vm[i, d] -= 1e-6
        num = (loss(vp) - loss(vm)) / 2e-6
        assert abs(num - g[i, d]) < 1e-5 * max(1.0, abs(num))
    # hard limit: with a tiny sigma the silhouette is the rasteriser's coverage
    hard, _ = O.soft_silhouette(v, f, cam, 32, 48, sigma=1e-7, blur_radius=0.0)
    pix, _, _, _ = O.rasterize(v.astype(np.float32), f, cam, 32, 48)
    assert ((hard > 0.5) == (pix >= 0)).mean() > 0.995


def test_rasteriser_oracle_agrees_with_independent_ray_casting():
    """pytorch3d is absent, so the oracle cannot be pinned to it; as an independent derivation, cast a ray through every
    pixel centre (float64 Moller-Trumbore against the view-space triangles): the nearest hit must be the oracle's face, its
    3-D barycentrics the oracle's perspective-corrected ones, its depth the oracle's zbuf."""
    v, f = S.make_test_mesh("blob", n_lat=10, n_lon=16)
    v = O.normalize_mesh(v)
    H = W = 64
    for params in (VIEWS[0], VIEWS[2], (2.0, 315.0, 135.0, 0.0, 0.3)):
        cam = O.camera(params)
        pix, bary, zbuf, _ = O.rasterize(v, f, cam, H, W)
        view = v.astype(np.float64) @ cam["R"].astype(np.float64) + cam["T"].astype(np.float64)
        tri = view[f]                                              # [Nf,3,3]
        ys, xs = O.pixel_ndc(H, W, np.float64), O.pixel_ndc(W, H, np.float64)
        s = float(cam["s"])
        d = np.stack(np.broadcast_arrays(xs[None, :] / s, ys[:, None] / s, np.ones((H, W))), -1).reshape(-1, 3)   # rays from the origin
        e1, e2 = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
        pv = np.cross(d[:, None, :], e2[None])                     # [P,Nf,3]
        det = (pv * e1[None]).sum(-1)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            tv = -tri[None, :, 0, :]                                # origin - v0
            u = (tv * pv).sum(-1) * inv
            qv = np.cross(tv, e1[None])
            w = (d[:, None, :] * qv).sum(-1) * inv
            t = (e2[None] * qv).sum(-1) * inv
        hit = (np.abs(det) > 1e-14) & (u > 0) & (w > 0) & (u + w < 1) & (t > 0)
        t_hit = np.where(hit, t, np.inf)
        best = t_hit.argmin(1)
        any_hit = np.isfinite(t_hit.min(1))
        ray_face = np.where(any_hit, best, -1).reshape(H, W)
        agree = ray_face == pix
        assert agree.mean() > 0.995, agree.mean()                   # pixel centres on an edge may go either way
        sel = agree & (pix >= 0)
        idx = np.nonzero(sel.reshape(-1))[0]
        bu, bw = u[idx, best[idx]], w[idx, best[idx]]
        ray_bary = np.stack([1 - bu - bw, bu, bw], -1)
        assert np.abs(ray_bary - bary[sel]).max() < 2e-4
        assert np.abs(t[idx, best[idx]] - zbuf[sel]).max() < 2e-5   # ray direction has unit z: t is the view-space depth


def test_phong_oracle_analytic_pixel():
    """HardPhong with one point light on a fronto-parallel triangle: at the pixel on the optical axis the light, view and
    normal directions coincide, so colour = (ambient + diffuse) * texture + specular (front side) or ambient * texture (the
    lighting is one-sided: a triangle wound the other way gets neither diffuse nor specular light)."""
    cam = O.camera((2.0, 0, 0, 0, 0))
    v = np.array([[-0.5, -0.5, 0], [0.5, -0.5, 0], [0, 0.6, 0]], dtype=np.float32)
    col = np.full((3, 3), 0.5, np.float32)
    H = 65                                                           # odd: pixel (32, 32) is on the optical axis
    front = O.render_phong(v, np.array([[0, 1, 2]]), col, cam, [0, 0, 3], H, H)   # normal +z, towards camera and light
    back = O.render_phong(v, np.array([[0, 2, 1]]), col, cam, [0, 0, 3], H, H)
    assert (front[0, 0] == 255).all() and (back[0, 0] == 255).all()               # white background
    assert (front[32, 32] == int((0.5 + 0.3) * 0.5 * 255 + 0.2 * 255)).all()      # 153
    assert (back[32, 32] == int(0.5 * 0.5 * 255)).all()                           # 63
    n = O.vertex_normals(v, np.array([[0, 1, 2]]))
    assert np.allclose(n, [[0, 0, 1]] * 3, atol=1e-6)


def test_point_rasteriser_oracle_picks_the_nearest_point_in_its_disc():
    from oracle import raster as O

    cam = O.camera((2.0, 0.0, 0.0, 0.0, 0.0))
    pts = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.5], [0.3, 0.2, 0.0]], np.float32)   # the second point is nearer to the camera at z = 2
    m = O.rasterize_points(pts, cam, 64, 64, 0.05)
    pr = O.project(pts, cam)
    assert pr[1, 2] < pr[0, 2]
    centre = m[31:33, 31:33]
    assert np.all(centre == 1)                     # same pixel, smaller depth wins
    assert (m == 2).sum() > 0 and (m == 0).sum() >= 0 and (m >= 0).sum() < 64 * 64 * 0.05
    xs = O.pixel_ndc(64, 64)
    yy, xx = np.nonzero(m == 2)
    assert np.all((xs[xx] - pr[2, 0]) ** 2 + (xs[yy] - pr[2, 1]) ** 2 < 0.05 ** 2)
