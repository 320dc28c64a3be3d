"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float32 arithmetic, no FMA contraction) of the "Render" step of
Render-Localise-Lift: the reference's `get_rasterizer` + `project_vertices_and_create_mask`
(preprocess_data/render_mesh_utils.py:115-174) and `normalize_mesh` (utils/demo_utils.py:130-145), which run
pytorch3d's `look_at_view_transform`, `FoVPerspectiveCameras` and `MeshRasterizer` (blur_radius 0, faces_per_pixel 1).

PARITY UNPINNED: pytorch3d (`git+https://github.com/facebookresearch/pytorch3d.git@stable`, requirements.txt:28, no
version pin) is neither under /root/reference nor installed here, and the reference holds no test or golden vector for
this step.  What follows restates pytorch3d's published algorithm:
  * look_at_view_transform(dist, elev, azim) in degrees: camera centre C = dist*(cos(el)sin(az), sin(el), cos(el)cos(az)),
    z = normalize(at - C), x = normalize(up x z), y = normalize(z x x) (x re-derived from y x z when degenerate), R has
    the axes as COLUMNS, T = -R^T C, row-vector convention X_view = X_world R + T;
  * FoVPerspectiveCameras defaults fov 60 deg, aspect 1, znear 1, zfar 100: x_ndc = x / (z tan(fov/2)), same for y; the
    rasteriser keeps view-space z as depth (MeshRasterizer.transform);
  * NDC has +X left and +Y up; pixel (row i, col j) samples y = 1 - (2i+1)/H, x = 1 - (2j+1)/W;
  * per pixel and face: signed-area barycentrics with area + 1e-8, perspective correction w_k' ~ w_k * prod(z_other)
    / max(sum, 1e-8), depth pz = sum w_k' z_k, a pixel is covered when all three corrected coordinates are > 0;
    faces with |area| <= 1e-8, max z < 0 or pz < 0 are skipped; nearest pz wins, ties go to the lower face index;
  * z_clip_value = znear / 2 for perspective cameras: triangles entirely behind it are culled; pytorch3d clips triangles
    that straddle the plane into sub-triangles -- restated here as "covered only where pz >= z_clip" (identical coverage
    while every vertex has z > 0; faces with a vertex at z <= 0 that straddle the plane are skipped and counted).
Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module."""
from __future__ import annotations

import numpy as np

F = np.float32
K_EPS = F(1e-8)


def normalize_mesh(verts, scale_factor=1.0):
    """utils/demo_utils.py:130-145."""
    v = np.asarray(verts, dtype=F)
    c = v - v.mean(0, dtype=F)
    size = (c.max(0) - c.min(0)).max()
    return (c / size * F(scale_factor)).astype(F)


def look_at_view_transform(dist, elev, azim):
    """pytorch3d.renderer.cameras.look_at_view_transform (degrees, at = origin, up = +Y) -> R [3,3], T [3], C [3]."""
    d, el, az = F(dist), F(np.deg2rad(F(elev))), F(np.deg2rad(F(azim)))
    C = np.array([d * np.cos(el) * np.sin(az), d * np.sin(el), d * np.cos(el) * np.cos(az)], dtype=F)

    def nrm(v):
        return (v / max(np.sqrt((v * v).sum(dtype=F)), F(1e-5))).astype(F)  # F.normalize eps = 1e-5 in look_at_rotation

    up = np.array([0, 1, 0], dtype=F)
    z = nrm(-C)
    x = nrm(np.cross(up, z).astype(F))
    y = nrm(np.cross(z, x).astype(F))
    if np.all(np.abs(x) <= 5e-3):
        x = nrm(np.cross(y, z).astype(F))
    R = np.stack([x, y, z], axis=1).astype(F)  # axes as columns
    T = (-(R.T @ C)).astype(F)
    return R, T, C


def camera(cam_params, fov_deg=60.0, znear=1.0):
    """render_mesh_utils.py:115-121: (distance, elevation, azimuth, x_trans, y_trans) -> dict(R, T, C, s, z_clip)."""
    dist, elev, azim, xt, yt = [float(c) for c in cam_params]
    R, T, C = look_at_view_transform(dist, elev, azim)
    T = T.copy()
    T[1] += F(yt)
    T[0] += F(xt)
    s = F(1.0) / F(np.tan(F(np.deg2rad(F(fov_deg))) / F(2)))
    return dict(R=R, T=T, C=C, s=F(s), z_clip=F(znear / 2.0))


def project(verts, cam):
    """[Nv,3] world -> [Nv,3] (x_ndc, y_ndc, z_view), float32, fixed operation order (matches the CUDA kernel)."""
    v = np.asarray(verts, dtype=F)
    R, T = cam["R"], cam["T"]
    out = np.empty_like(v)
    view = np.empty_like(v)
    for k in range(3):
        view[:, k] = ((v[:, 0] * R[0, k] + v[:, 1] * R[1, k]) + v[:, 2] * R[2, k]) + T[k]
    fx, fy = F(cam.get("fx", cam.get("s", 1.0))), F(cam.get("fy", cam.get("s", 1.0)))
    out[:, 0] = (view[:, 0] * fx) / view[:, 2] + F(cam.get("cx", 0.0))
    out[:, 1] = (view[:, 1] * fy) / view[:, 2] + F(cam.get("cy", 0.0))
    out[:, 2] = view[:, 2]
    return out


def pixel_ndc(S1, S2, dtype=F):
    """pytorch3d PixToNonSquareNdc on the flipped index for every pixel of a side of S1 pixels (other side S2): the longer
    side spans [-S1/S2, S1/S2], the shorter [-1, 1]."""
    rng = dtype(S1) / dtype(S2) * dtype(2) if S1 > S2 else dtype(2)
    off = rng / dtype(2)
    return (-off + (rng * (S1 - 1 - np.arange(S1)).astype(dtype) + off) / dtype(S1)).astype(dtype)


def perspective_camera(focal_length, principal_point, image_size):
    """optim/renderer.py:29-46: pytorch3d PerspectiveCameras(in_ndc=False) with R = diag(-1,-1,1), T = 0; focal length and
    principal point in pixels are converted to NDC with s = min(H, W) / 2: f_ndc = f / s, c_ndc = -(c - size/2) / s.
    No znear -> no z clipping."""
    H, W = image_size
    fl = np.broadcast_to(np.asarray(focal_length, dtype=np.float64).reshape(-1), (2,)) if np.ndim(focal_length) else np.array([focal_length] * 2, dtype=np.float64)
    pp = np.asarray(principal_point, dtype=np.float64).reshape(2)
    s = min(H, W) / 2.0
    return dict(R=np.diag([-1.0, -1.0, 1.0]).astype(F), T=np.zeros(3, F), C=np.zeros(3, F), fx=F(fl[0] / s), fy=F(fl[1] / s),
                cx=F(-(pp[0] - W / 2.0) / s), cy=F(-(pp[1] - H / 2.0) / s), z_clip=F(0.0))


def _edge(px, py, ax, ay, bx, by):
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax)


def rasterize(verts, faces, cam, H, W):
    """pytorch3d rasterize_meshes (naive path), K=1, blur 0, perspective_correct.  Returns pix_to_face [H,W] int32 (-1 =
    background), bary [H,W,3] f32 (-1 on background, like pytorch3d), zbuf [H,W] f32 (-1 on background), n_skipped."""
    with np.errstate(divide="ignore", invalid="ignore"):
        p = project(verts, cam)
    faces = np.asarray(faces, dtype=np.int64)
    pix = np.full((H, W), -1, dtype=np.int32)
    bary = np.full((H, W, 3), -1, dtype=F)
    zbuf = np.full((H, W), -1, dtype=F)
    best = np.full((H, W), np.inf, dtype=F)
    ys, xs = pixel_ndc(H, W), pixel_ndc(W, H)
    zc = cam["z_clip"]
    skipped = 0
    for f, (i0, i1, i2) in enumerate(faces):
        (x0, y0, z0), (x1, y1, z1), (x2, y2, z2) = p[i0], p[i1], p[i2]
        zmin, zmax = min(z0, z1, z2), max(z0, z1, z2)
        if zmax < max(zc, F(0)):
            continue  # entirely behind the clip plane (clip_faces case 1) / behind the camera
        if zmin <= 0:
            skipped += 1  # straddles the camera plane: pytorch3d clips it; not restated
            continue
        area = _edge(x2, y2, x0, y0, x1, y1)
        if -K_EPS <= area <= K_EPS:
            continue
        xmin, xmax, ymin, ymax = min(x0, x1, x2), max(x0, x1, x2), min(y0, y1, y2), max(y0, y1, y2)
        # pixel rows/cols whose centre lies inside the bounding box (x decreases with the column index)
        cols = np.nonzero((xs >= xmin) & (xs <= xmax))[0]
        rows = np.nonzero((ys >= ymin) & (ys <= ymax))[0]
        if len(cols) == 0 or len(rows) == 0:
            continue
        px = xs[cols][None, :]
        py = ys[rows][:, None]
        den = area + K_EPS
        w0 = _edge(px, py, x1, y1, x2, y2) / den
        w1 = _edge(px, py, x2, y2, x0, y0) / den
        w2 = _edge(px, py, x0, y0, x1, y1) / den
        t0 = (w0 * z1) * z2
        t1 = (z0 * w1) * z2
        t2 = (z0 * z1) * w2
        d = np.maximum((t0 + t1) + t2, K_EPS)
        b0, b1, b2 = t0 / d, t1 / d, t2 / d
        pz = (b0 * z0 + b1 * z1) + b2 * z2
        inside = (b0 > 0) & (b1 > 0) & (b2 > 0) & (pz >= 0)
        if zc > 0 and zmin < zc:
            inside &= pz >= zc  # the part of a straddling triangle that pytorch3d's clip_faces keeps
        r = rows[:, None].repeat(len(cols), 1)
        c = cols[None, :].repeat(len(rows), 0)
        cur = best[r, c]
        take = inside & (pz < cur)  # faces visited in ascending order: ties keep the lower index
        rr, cc = r[take], c[take]
        best[rr, cc] = pz[take]
        pix[rr, cc] = f
        zbuf[rr, cc] = pz[take]
        bary[rr, cc, 0], bary[rr, cc, 1], bary[rr, cc, 2] = b0[take], b1[take], b2[take]
    return pix, bary, zbuf, skipped


def project_vertices_and_create_mask(verts, faces, cam_params, contact_vertices, image_size, min_vertices=3):
    """render_mesh_utils.py:123-174 -> (mask uint8 [H,W] in {0,255}, pixel_to_vertices_map int64 [H,W,3] (-1 background),
    bary f32 [H,W,3])."""
    H, W = image_size
    faces = np.asarray(faces, dtype=np.int64)
    pix, bary, _, _ = rasterize(verts, faces, camera(cam_params), H, W)
    cnt = np.isin(faces, list(set(int(v) for v in contact_vertices))).sum(1)
    hot = cnt >= min_vertices
    mask = ((pix >= 0) & hot[np.maximum(pix, 0)]).astype(np.uint8) * 255
    p2v = np.full((H, W, 3), -1, dtype=np.int64)
    p2v[pix >= 0] = faces[pix[pix >= 0]]
    return mask, p2v, bary


def vertex_normals(verts, faces):
    """pytorch3d Meshes.verts_normals_packed(): area-weighted sum of the face normals, normalised (eps 1e-6)."""
    v = np.asarray(verts, dtype=F)
    f = np.asarray(faces, dtype=np.int64)
    n = np.zeros_like(v)
    v0, v1, v2 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    np.add.at(n, f[:, 1], np.cross(v2 - v1, v0 - v1))
    np.add.at(n, f[:, 2], np.cross(v0 - v2, v1 - v2))
    np.add.at(n, f[:, 0], np.cross(v1 - v0, v2 - v0))
    return (n / np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-6)).astype(F)


def render_phong(verts, faces, colors, cam, light_location, H, W, ambient=0.5, diffuse=0.3, specular=0.2, shininess=64.0):
    """render_mesh (render_mesh_utils.py:177-198): MeshRenderer(MeshRasterizer, HardPhongShader) with one PointLights
    (ambient .5, diffuse .3, specular .2), default Materials (shininess 64), white background, `(image*255).astype(uint8)`.
    Returns uint8 [H,W,3]."""
    v = np.asarray(verts, dtype=F)
    f = np.asarray(faces, dtype=np.int64)
    col = np.asarray(colors, dtype=F)
    pix, bary, _, _ = rasterize(v, f, cam, H, W)
    nrm = vertex_normals(v, f)
    fg = pix >= 0
    fi = f[pix[fg]]
    b = bary[fg]
    interp = lambda a: (b[:, 0:1] * a[fi[:, 0]] + b[:, 1:2] * a[fi[:, 1]]) + b[:, 2:3] * a[fi[:, 2]]
    pts, nn, tex = interp(v), interp(nrm), interp(col)
    unit = lambda a: a / np.maximum(np.linalg.norm(a, axis=1, keepdims=True), 1e-6)
    nn = unit(nn)
    L = unit(np.asarray(light_location, dtype=F)[None] - pts)
    cosang = (nn * L).sum(1)
    diff = F(diffuse) * np.maximum(cosang, 0)
    view = unit(cam["C"][None] - pts)
    refl = -L + 2 * (cosang[:, None] * nn)
    alpha = np.maximum((view * refl).sum(1), 0) * (cosang > 0)
    spec = F(specular) * np.power(alpha, F(shininess))
    rgb = (F(ambient) + diff)[:, None] * tex + spec[:, None]
    img = np.ones((H, W, 3), dtype=F)
    img[fg] = rgb
    return (img * 255).astype(np.uint8)


# ------------------------------------------------------------------------------------------------ soft silhouette (optim/)
def _seg_dist2(px, py, ax, ay, bx, by, eps=1e-8):
    """pytorch3d PointLineDistanceForward: squared distance from p to the segment a-b (+ the clamped parameter)."""
    dx, dy = bx - ax, by - ay
    l2 = dx * dx + dy * dy
    if l2 <= eps:
        return (px - bx) ** 2 + (py - by) ** 2, 1.0
    t = min(max((dx * (px - ax) + dy * (py - ay)) / l2, 0.0), 1.0)
    qx, qy = ax + t * dx, ay + t * dy
    return (px - qx) ** 2 + (py - qy) ** 2, t


def soft_silhouette(verts, faces, cam, H, W, sigma=1e-4, blur_radius=None, K=100, grad_alpha=None):
    """SSRenderer.render (optim/renderer.py:64-104): pytorch3d MeshRasterizer(blur_radius = log(1/1e-4 - 1) sigma,
    faces_per_pixel = 100, perspective-correct, clipped barycentrics) + SoftSilhouetteShader(sigma):
      per pixel, every face whose squared NDC distance to the pixel centre is < blur_radius (or that covers it) is a candidate
      with signed distance d (negative inside); the K candidates nearest in depth are kept;
      alpha = 1 - prod_k (1 - sigmoid(-d_k / sigma)).
    float64, brute force.  Returns alpha [H,W], zbuf0 [H,W] (depth of the nearest kept candidate, -1 if none) and, when
    `grad_alpha` [H,W] is given, d(sum(grad_alpha * alpha)) / d verts [Nv,3] (pytorch3d's backward: through the distance to
    the closest edge with the clamped segment parameter held fixed, through the projection, not through depth ordering).
    PARITY UNPINNED (pytorch3d absent), see the module header."""
    v = np.asarray(verts, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int64)
    if blur_radius is None:
        blur_radius = np.log(1.0 / 1e-4 - 1.0) * sigma
    R, T = np.asarray(cam["R"], np.float64), np.asarray(cam["T"], np.float64)
    fx, fy = float(cam.get("fx", cam.get("s", 1.0))), float(cam.get("fy", cam.get("s", 1.0)))
    cx, cy = float(cam.get("cx", 0.0)), float(cam.get("cy", 0.0))
    view = v @ R + T
    ndc = np.stack([fx * view[:, 0] / view[:, 2] + cx, fy * view[:, 1] / view[:, 2] + cy, view[:, 2]], -1)
    ys, xs = pixel_ndc(H, W, np.float64), pixel_ndc(W, H, np.float64)
    alpha = np.zeros((H, W))
    zbuf0 = np.full((H, W), -1.0)
    g_ndc = np.zeros((len(v), 2))
    pad = np.sqrt(blur_radius)
    tri = ndc[f]                                   # [Nf,3,3]
    lo, hi = tri[:, :, :2].min(1) - pad, tri[:, :, :2].max(1) + pad
    zmax, zmin = tri[:, :, 2].max(1), tri[:, :, 2].min(1)
    for r in range(H):
        py = ys[r]
        row_faces = np.nonzero((lo[:, 1] <= py) & (hi[:, 1] >= py) & (zmax >= 0) & (zmin > 0))[0]
        for c in range(W):
            px = xs[c]
            cand = []
            for fi in row_faces[(lo[row_faces, 0] <= px) & (hi[row_faces, 0] >= px)]:
                (x0, y0, z0), (x1, y1, z1), (x2, y2, z2) = tri[fi]
                area = _edge(x2, y2, x0, y0, x1, y1)
                if abs(area) <= 1e-8:
                    continue
                den = area + 1e-8
                w = np.array([_edge(px, py, x1, y1, x2, y2), _edge(px, py, x2, y2, x0, y0), _edge(px, py, x0, y0, x1, y1)]) / den
                top = np.array([w[0] * z1 * z2, z0 * w[1] * z2, z0 * z1 * w[2]])
                b = top / max(top.sum(), 1e-8)
                bc = np.clip(b, 0.0, 1.0)
                bc = bc / max(bc.sum(), 1e-5)
                pz = bc[0] * z0 + bc[1] * z1 + bc[2] * z2
                if pz < 0:
                    continue
                segs = [_seg_dist2(px, py, x0, y0, x1, y1), _seg_dist2(px, py, x1, y1, x2, y2), _seg_dist2(px, py, x2, y2, x0, y0)]
                e = int(np.argmin([s_[0] for s_ in segs]))
                dist = segs[e][0]
                inside = bool((b > 0).all())
                if not inside and dist >= blur_radius:
                    continue
                cand.append((pz, fi, -dist if inside else dist, e, segs[e][1]))
            if not cand:
                continue
            cand.sort(key=lambda t: (t[0], t[1]))
            cand = cand[:K]
            prob = np.array([1.0 / (1.0 + np.exp(c_[2] / sigma)) if c_[2] / sigma < 700 else 0.0 for c_ in cand])
            keep = np.prod(1.0 - prob)
            alpha[r, c] = 1.0 - keep
            zbuf0[r, c] = cand[0][0]
            if grad_alpha is not None and grad_alpha[r, c] != 0.0:
                for (pz, fi, sd, e, t), p_k in zip(cand, prob):
                    # d alpha / d sd = -(prod_{j != k}(1 - p_j)) p_k (1 - p_k) / sigma
                    others = np.prod([1.0 - q for j, q in enumerate(prob) if not (cand[j][1] == fi)]) if p_k >= 1.0 else keep / (1.0 - p_k)
                    g_sd = -grad_alpha[r, c] * others * p_k * (1.0 - p_k) / sigma
                    g_d = g_sd * (-1.0 if sd < 0 else 1.0)
                    ia, ib = f[fi][e], f[fi][(e + 1) % 3]
                    ax, ay = ndc[ia, 0], ndc[ia, 1]
                    bx, by = ndc[ib, 0], ndc[ib, 1]
                    qx, qy = ax + t * (bx - ax), ay + t * (by - ay)
                    gq = g_d * 2.0 * np.array([qx - px, qy - py])       # d dist / d q, the closest point on the segment
                    g_ndc[ia] += (1.0 - t) * gq
                    g_ndc[ib] += t * gq
    if grad_alpha is None:
        return alpha, zbuf0
    g_view = np.zeros_like(v)
    z = view[:, 2]
    g_view[:, 0] = g_ndc[:, 0] * fx / z
    g_view[:, 1] = g_ndc[:, 1] * fy / z
    g_view[:, 2] = -(g_ndc[:, 0] * fx * view[:, 0] + g_ndc[:, 1] * fy * view[:, 1]) / (z * z)
    return alpha, zbuf0, g_view @ R.T


def rasterize_points(points, cam, H, W, radius):
    """pytorch3d PointsRasterizer as utils_obj_pc.py:88-113 reads it (num_point2pixel == 1) -> pixel_to_point int64 [H,W]
    (-1 background): per pixel the point nearest in depth with (x_pix - x)^2 + (y_pix - y)^2 < radius^2 and z >= 0, equal
    depths to the lower index.  float32 with the operation order of the CUDA kernel.  Parity unpinned like the rest of this
    file (pytorch3d absent): conventions as documented for ivlm_rasterize_mesh."""
    pr = project(points, cam)
    xs, ys = pixel_ndc(W, H), pixel_ndc(H, W)
    best_z = np.full((H, W), np.inf, dtype=F)
    best_i = np.full((H, W), -1, dtype=np.int64)
    r2 = F(radius) * F(radius)
    for i in range(pr.shape[0]):
        x, y, z = pr[i]
        if not z >= 0:
            continue
        dx = (xs - x).astype(F)
        dy = (ys - y).astype(F)
        d2 = (dx * dx)[None, :] + (dy * dy)[:, None]
        hit = (d2 < r2) & (z < best_z)          # ascending index order: a later point only wins with a strictly smaller depth
        best_z[hit] = z
        best_i[hit] = i
    return best_i
