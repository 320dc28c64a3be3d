"""Seeded synthetic data of the reference's shapes (no datasets or checkpoints exist offline).

Used by bench.py, the tests and oracle/make_goldens.py so that all of them see identical inputs.
Only numpy here (np.random.default_rng / PCG64 is reproducible across machines).
"""
from __future__ import annotations

import numpy as np

HUMAN_VIEWS = ("topfront", "bottomfront", "topback", "bottomback")  # preprocess_data/constants.py:315-382
N_SMPL = 6890
N_SMPLX = 10475
# normalize_cam_params(HUMAN_VIEW_DICT['4MV-Z_Vitru']['cam_params']) -- datasets/base_contact_dataset.py:37-50
HCONTACT_CAM_PARAMS = np.array([[.2, .125, .875, .5, .5], [.2, .875, .875, .5, .65],
                                [.2, .125, .375, .5, .5], [.2, .875, .375, .5, .65]], dtype=np.float32)


def _silhouette(rng, size: int, coverage: float) -> np.ndarray:
    """Union of ellipses covering roughly `coverage` of the image."""
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    mask = np.zeros((size, size), bool)
    target = coverage * size * size
    tries = 0
    while mask.sum() < target and tries < 64:
        cy, cx = rng.uniform(0.25, 0.75, 2) * size
        ry, rx = rng.uniform(0.05, 0.22, 2) * size
        mask |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
        tries += 1
    return mask


def make_mesh_lift_maps(n_verts: int = N_SMPL, n_views: int = 4, size: int = 1024, coverage: float = 0.18,
                        tile: int = 8, seed: int = 0):
    """pixel_to_vertex [V,size,size,3] int64 (-1 background) and bary [V,size,size,3] float32.

    Spatially coherent like a rasterised mesh: the image is cut into `tile`-pixel cells, every cell is two
    triangles over a lattice of vertices, neighbouring cells share vertices, barycentrics are the true
    coordinates of the pixel centre inside its triangle (SURVEY.md section 8d explains why i.i.d. random
    triangles would collapse every probability to ~0.5).
    """
    rng = np.random.default_rng(seed)
    p2v = np.full((n_views, size, size, 3), -1, np.int64)
    bary = np.zeros((n_views, size, size, 3), np.float32)
    nl = size // tile + 1
    yy, xx = np.mgrid[0:size, 0:size]
    ty, tx = yy // tile, xx // tile
    fy = ((yy % tile) + 0.5) / tile
    fx = ((xx % tile) + 0.5) / tile
    upper = (fx + fy) <= 1.0
    for v in range(n_views):
        sil = _silhouette(rng, size, coverage)
        lattice = rng.integers(0, n_verts, size=(nl, nl))
        v00, v01 = lattice[ty, tx], lattice[ty, tx + 1]
        v10, v11 = lattice[ty + 1, tx], lattice[ty + 1, tx + 1]
        # upper triangle (v00, v01, v10): weights (1-fx-fy, fx, fy); lower (v11, v10, v01): (fx+fy-1, 1-fx, 1-fy)
        a = np.where(upper, v00, v11)
        b = np.where(upper, v01, v10)
        c = np.where(upper, v10, v01)
        wa = np.where(upper, 1.0 - fx - fy, fx + fy - 1.0)
        wb = np.where(upper, fx, 1.0 - fx)
        wc = np.where(upper, fy, 1.0 - fy)
        tri = np.stack([a, b, c], -1)
        w = np.stack([wa, wb, wc], -1).astype(np.float32)
        p2v[v][sil] = tri[sil]
        bary[v][sil] = w[sil]
    return p2v, bary


def make_point_lift_maps(n_points: int = 2048, n_views: int = 4, size: int = 1024, coverage: float = 0.12,
                         seed: int = 0) -> np.ndarray:
    """pixel_to_point [V,size,size] int64, -1 background (p2pmap_*.npz['mapping'], components.py:309-327)."""
    rng = np.random.default_rng(seed + 1000)
    out = np.full((n_views, size, size), -1, np.int64)
    cell = 8
    nl = size // cell
    yy, xx = np.mgrid[0:size, 0:size]
    for v in range(n_views):
        sil = _silhouette(rng, size, coverage)
        lattice = rng.integers(0, n_points, size=(nl, nl))
        ids = lattice[yy // cell, xx // cell]
        out[v][sil] = ids[sil]
    return out


def make_mask_logits(batch: int, n_views: int = 4, size: int = 1024, seed: int = 0, std: float = 4.0) -> np.ndarray:
    """Smooth random logits [B,V,size,size] float32 (bilinear blow-up of a coarse random field)."""
    rng = np.random.default_rng(seed + 2000)
    coarse = rng.normal(0.0, std, size=(batch, n_views, size // 32 + 1, size // 32 + 1)).astype(np.float32)
    yi = np.linspace(0, coarse.shape[2] - 1.001, size, dtype=np.float32)
    y0 = yi.astype(np.int64)
    fy = (yi - y0)[None, None, :, None]
    rows = coarse[:, :, y0, :] * (1 - fy) + coarse[:, :, y0 + 1, :] * fy
    fx = (yi - y0)[None, None, None, :]
    out = rows[:, :, :, y0] * (1 - fx) + rows[:, :, :, y0 + 1] * fx
    return np.ascontiguousarray(out, dtype=np.float32)


def make_smplx_matrix(n_smplx: int = N_SMPLX, n_smpl: int = N_SMPL, seed: int = 0) -> np.ndarray:
    """Dense [n_smplx, n_smpl] barycentric-style mapping with 3 non-zeros per row summing to 1
    (the shape of SMPL_TO_SMPLX_MAPPING's 'matrix', utils/utils.py:428-443)."""
    rng = np.random.default_rng(seed + 3000)
    m = np.zeros((n_smplx, n_smpl), np.float32)
    cols = rng.integers(0, n_smpl, size=(n_smplx, 3))
    w = rng.dirichlet((1.0, 1.0, 1.0), size=n_smplx).astype(np.float32)
    for k in range(3):
        np.add.at(m, (np.arange(n_smplx), cols[:, k]), w[:, k])
    return m


# ---------------------------------------------------------------------------------------------- model weights
def make_test_mesh(kind: str = "blob", n_lat: int = 24, n_lon: int = 48, seed: int = 0):
    """Closed triangle meshes for the rasteriser tests / bench: `blob` = UV sphere with smooth radial noise (self-occluding
    from most views), `torus`, `adversarial` = blob + coplanar overlapping quads (depth ties), degenerate faces, a face
    behind the camera plane, a large face crossing the clip plane and faces leaving the view.  -> (verts f32 [Nv,3], faces
    i64 [Nf,3]); extents ~[-0.5, 0.5] like normalize_mesh output."""
    rng = np.random.default_rng(seed)
    th = np.linspace(0, np.pi, n_lat + 1)[1:-1]
    ph = np.linspace(0, 2 * np.pi, n_lon, endpoint=False)
    T, Pp = np.meshgrid(th, ph, indexing="ij")
    if kind == "torus":
        th = np.linspace(0, 2 * np.pi, n_lat, endpoint=False)
        T, Pp = np.meshgrid(th, ph, indexing="ij")
        r = 0.33 + 0.14 * np.cos(T)
        grid = np.stack([r * np.cos(Pp), 0.14 * np.sin(T), r * np.sin(Pp)], -1)
        verts = grid.reshape(-1, 3)
        idx = np.arange(n_lat * n_lon).reshape(n_lat, n_lon)
        a, b = idx, np.roll(idx, -1, 1)
        c, d = np.roll(idx, -1, 0), np.roll(np.roll(idx, -1, 0), -1, 1)
        faces = np.concatenate([np.stack([a, c, b], -1).reshape(-1, 3), np.stack([b, c, d], -1).reshape(-1, 3)])
        return verts.astype(np.float32), faces.astype(np.int64)
    k = rng.normal(size=(4, 3))
    dirs = np.stack([np.sin(T) * np.cos(Pp), np.cos(T), np.sin(T) * np.sin(Pp)], -1)
    rad = 0.38 + 0.1 * np.sin(dirs @ k[0] * 3) * np.cos(dirs @ k[1] * 2) + 0.04 * np.sin(dirs @ k[2] * 7)
    body = (dirs * rad[..., None]).reshape(-1, 3)
    north, south = np.array([[0, 0.4, 0]]), np.array([[0, -0.4, 0]])
    verts = np.concatenate([body, north, south])
    n_body = body.shape[0]
    idx = np.arange(n_body).reshape(n_lat - 1, n_lon)
    a, b = idx[:-1], np.roll(idx[:-1], -1, 1)
    c, d = idx[1:], np.roll(idx[1:], -1, 1)
    faces = [np.stack([a, b, c], -1).reshape(-1, 3), np.stack([b, d, c], -1).reshape(-1, 3)]
    top, bot = idx[0], idx[-1]
    faces.append(np.stack([np.full(n_lon, n_body), np.roll(top, -1), top], -1))
    faces.append(np.stack([np.full(n_lon, n_body + 1), bot, np.roll(bot, -1)], -1))
    faces = np.concatenate(faces)
    if kind == "adversarial":
        n0 = verts.shape[0]
        extra = np.array([
            [-0.3, -0.3, 0.45], [0.3, -0.3, 0.45], [0.3, 0.3, 0.45], [-0.3, 0.3, 0.45],      # quad A (z = 0.45)
            [-0.2, -0.35, 0.45], [0.35, -0.2, 0.45], [0.2, 0.35, 0.45], [-0.35, 0.2, 0.45],   # quad B, coplanar with A
            [0.1, 0.1, 0.1], [0.1, 0.1, 0.1], [0.2, 0.2, 0.2],                                  # degenerate (two equal verts)
            [0.0, 0.0, 5.0], [0.1, 0.0, 5.0], [0.0, 0.1, 5.0],                                  # behind a camera at z = 2
            [-0.6, -0.2, 0.2], [0.6, -0.2, 0.2], [0.0, 0.1, 1.7],                               # crosses z_clip for that camera
            [0.4, 0.4, 0.0], [3.0, 0.5, 0.0], [0.5, 3.0, 0.0],                                  # leaves the view
        ])
        verts = np.concatenate([verts, extra])
        ef = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7], [8, 9, 10], [11, 12, 13], [14, 15, 16], [17, 18, 19]]) + n0
        faces = np.concatenate([faces, ef, ef[:2]])  # the last two duplicate quad A exactly (identical depth everywhere)
    return verts.astype(np.float32), faces.astype(np.int64)


CLIP_PREFIX = "model.vision_tower.vision_tower.vision_model."
SAM_PREFIX = "model.visual_model."


def state_dict_spec(cfg) -> dict:
    """name -> (shape, kind) for every tensor on the hot path, with the reference's checkpoint key names
    (probed from the reference model's state_dict()).
    kind: 'w' linear/conv weight (fan-in scaled normal), 'b' bias, 'g' norm gain (~1), 'e' embedding-like table,
    'pe' the prompt encoder's gaussian matrix."""
    s = {}
    D, F, Vv = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    s["model.embed_tokens.weight"] = ((Vv, D), "e")
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}."
        for n in ("q", "k", "v", "o"):
            s[p + f"self_attn.{n}_proj.weight"] = ((D, D), "w")
        s[p + "mlp.gate_proj.weight"] = ((F, D), "w")
        s[p + "mlp.up_proj.weight"] = ((F, D), "w")
        s[p + "mlp.down_proj.weight"] = ((D, F), "w")
        s[p + "input_layernorm.weight"] = ((D,), "g")
        s[p + "post_attention_layernorm.weight"] = ((D,), "g")
    s["model.norm.weight"] = ((D,), "g")
    s["lm_head.weight"] = ((Vv, D), "w")
    # CLIP (HF CLIPVisionModel)
    C, CF, ps = cfg.clip_hidden_size, cfg.clip_intermediate_size, cfg.clip_patch_size
    c = CLIP_PREFIX
    s[c + "embeddings.class_embedding"] = ((C,), "e")
    s[c + "embeddings.patch_embedding.weight"] = ((C, 3, ps, ps), "w")
    s[c + "embeddings.position_embedding.weight"] = ((cfg.clip_tokens, C), "e")
    s[c + "pre_layrnorm.weight"] = ((C,), "g")
    s[c + "pre_layrnorm.bias"] = ((C,), "b")
    for i in range(cfg.clip_num_hidden_layers):
        p = c + f"encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s[p + f"self_attn.{n}.weight"] = ((C, C), "w")
            s[p + f"self_attn.{n}.bias"] = ((C,), "b")
        s[p + "layer_norm1.weight"] = ((C,), "g"); s[p + "layer_norm1.bias"] = ((C,), "b")
        s[p + "mlp.fc1.weight"] = ((CF, C), "w"); s[p + "mlp.fc1.bias"] = ((CF,), "b")
        s[p + "mlp.fc2.weight"] = ((C, CF), "w"); s[p + "mlp.fc2.bias"] = ((C,), "b")
        s[p + "layer_norm2.weight"] = ((C,), "g"); s[p + "layer_norm2.bias"] = ((C,), "b")
    s[c + "post_layernorm.weight"] = ((C,), "g"); s[c + "post_layernorm.bias"] = ((C,), "b")
    s["model.mm_projector.weight"] = ((D, C), "w"); s["model.mm_projector.bias"] = ((D,), "b")
    # SAM image encoder
    E, G, hd = cfg.sam_embed_dim, cfg.sam_grid, cfg.sam_embed_dim // cfg.sam_num_heads
    e = SAM_PREFIX + "image_encoder."
    s[e + "pos_embed"] = ((1, G, G, E), "e")
    s[e + "patch_embed.proj.weight"] = ((E, 3, cfg.sam_patch_size, cfg.sam_patch_size), "w")
    s[e + "patch_embed.proj.bias"] = ((E,), "b")
    for i in range(cfg.sam_depth):
        p = e + f"blocks.{i}."
        sz = G if i in cfg.sam_global_attn_indexes else cfg.sam_window_size
        s[p + "norm1.weight"] = ((E,), "g"); s[p + "norm1.bias"] = ((E,), "b")
        s[p + "attn.rel_pos_h"] = ((2 * sz - 1, hd), "e"); s[p + "attn.rel_pos_w"] = ((2 * sz - 1, hd), "e")
        s[p + "attn.qkv.weight"] = ((3 * E, E), "w"); s[p + "attn.qkv.bias"] = ((3 * E,), "b")
        s[p + "attn.proj.weight"] = ((E, E), "w"); s[p + "attn.proj.bias"] = ((E,), "b")
        s[p + "norm2.weight"] = ((E,), "g"); s[p + "norm2.bias"] = ((E,), "b")
        s[p + "mlp.lin1.weight"] = ((4 * E, E), "w"); s[p + "mlp.lin1.bias"] = ((4 * E,), "b")
        s[p + "mlp.lin2.weight"] = ((E, 4 * E), "w"); s[p + "mlp.lin2.bias"] = ((E,), "b")
    O = cfg.sam_out_chans
    s[e + "neck.0.weight"] = ((O, E, 1, 1), "w")
    s[e + "neck.1.weight"] = ((O,), "g"); s[e + "neck.1.bias"] = ((O,), "b")
    s[e + "neck.2.weight"] = ((O, O, 3, 3), "w")
    s[e + "neck.3.weight"] = ((O,), "g"); s[e + "neck.3.bias"] = ((O,), "b")
    # prompt encoder (only the tensors the text-prompt path reads)
    pe = SAM_PREFIX + "prompt_encoder."
    s[pe + "pe_layer.positional_encoding_gaussian_matrix"] = ((2, O // 2), "pe")
    s[pe + "no_mask_embed.weight"] = ((1, O), "e")
    # mask decoder(s): the '*-DifDe' token types carry a human and an object copy next to the shared one (InteractVLM.py:114-122)
    nm = cfg.sam_num_multimask_outputs + 1
    decoders = ["mask_decoder"] + (["human_mask_decoder", "object_mask_decoder"] if "DifDe" in cfg.token_type else [])
    for _dec in decoders:
        d = SAM_PREFIX + _dec + "."

        def attn(p, internal):
            for n in ("q_proj", "k_proj", "v_proj"):
                s[p + n + ".weight"] = ((internal, O), "w"); s[p + n + ".bias"] = ((internal,), "b")
            s[p + "out_proj.weight"] = ((O, internal), "w"); s[p + "out_proj.bias"] = ((O,), "b")

        for i in range(cfg.sam_dec_depth):
            p = d + f"transformer.layers.{i}."
            attn(p + "self_attn.", O)
            attn(p + "cross_attn_token_to_image.", O // 2)
            attn(p + "cross_attn_image_to_token.", O // 2)
            for n in ("norm1", "norm2", "norm3", "norm4"):
                s[p + n + ".weight"] = ((O,), "g"); s[p + n + ".bias"] = ((O,), "b")
            s[p + "mlp.lin1.weight"] = ((cfg.sam_dec_mlp_dim, O), "w"); s[p + "mlp.lin1.bias"] = ((cfg.sam_dec_mlp_dim,), "b")
            s[p + "mlp.lin2.weight"] = ((O, cfg.sam_dec_mlp_dim), "w"); s[p + "mlp.lin2.bias"] = ((O,), "b")
        attn(d + "transformer.final_attn_token_to_image.", O // 2)
        s[d + "transformer.norm_final_attn.weight"] = ((O,), "g"); s[d + "transformer.norm_final_attn.bias"] = ((O,), "b")
        s[d + "iou_token.weight"] = ((1, O), "e")
        s[d + "mask_tokens.weight"] = ((nm, O), "e")
        s[d + "output_upscaling.0.weight"] = ((O, O // 4, 2, 2), "w"); s[d + "output_upscaling.0.bias"] = ((O // 4,), "b")
        s[d + "output_upscaling.1.weight"] = ((O // 4,), "g"); s[d + "output_upscaling.1.bias"] = ((O // 4,), "b")
        s[d + "output_upscaling.3.weight"] = ((O // 4, O // 8, 2, 2), "w"); s[d + "output_upscaling.3.bias"] = ((O // 8,), "b")
        for i in range(nm):
            p = d + f"output_hypernetworks_mlps.{i}.layers."
            s[p + "0.weight"] = ((O, O), "w"); s[p + "0.bias"] = ((O,), "b")
            s[p + "1.weight"] = ((O, O), "w"); s[p + "1.bias"] = ((O,), "b")
            s[p + "2.weight"] = ((O // 8, O), "w"); s[p + "2.bias"] = ((O // 8,), "b")
    # [SEG] projection and camera gate
    s["model.text_hidden_fcs.0.0.weight"] = ((D, D), "w"); s["model.text_hidden_fcs.0.0.bias"] = ((D,), "b")
    s["model.text_hidden_fcs.0.2.weight"] = ((cfg.out_dim, D), "w"); s["model.text_hidden_fcs.0.2.bias"] = ((cfg.out_dim,), "b")
    O_ = cfg.out_dim
    if cfg.multiview_cam_cond and cfg.cam_encoder_type == "vi_v1":       # components.py:541-572
        s["cam_pose_encoder.spatial_encoder.0.weight"] = ((128, 5), "w"); s["cam_pose_encoder.spatial_encoder.0.bias"] = ((128,), "b")
        s["cam_pose_encoder.spatial_encoder.2.weight"] = ((128, 128), "w"); s["cam_pose_encoder.spatial_encoder.2.bias"] = ((128,), "b")
        for v in range(cfg.multiview_channels):
            s[f"cam_pose_encoder.view_transforms.{v}.weight"] = ((O_, 128), "w")
            s[f"cam_pose_encoder.view_transforms.{v}.bias"] = ((O_,), "b")
    elif cfg.multiview_cam_cond and cfg.cam_encoder_type == "view_index":  # components.py:510-539
        s["cam_pose_encoder.spatial_encoder.0.weight"] = ((O_, 5), "w"); s["cam_pose_encoder.spatial_encoder.0.bias"] = ((O_,), "b")
        s["cam_pose_encoder.spatial_encoder.2.weight"] = ((O_, O_), "w"); s["cam_pose_encoder.spatial_encoder.2.bias"] = ((O_,), "b")
        for v in range(cfg.multiview_channels):
            s[f"cam_pose_encoder.view_transforms.{v}.weight"] = ((O_, O_), "w")
            s[f"cam_pose_encoder.view_transforms.{v}.bias"] = ((O_,), "b")
    elif cfg.multiview_cam_cond and cfg.cam_encoder_type == "simple":      # components.py:491-508
        s["cam_pose_encoder.linear1.weight"] = ((O_, 5), "w"); s["cam_pose_encoder.linear1.bias"] = ((O_,), "b")
    if cfg.token_type.replace("-DifDe", "") in ("Gen-Hu-Obj", "Gen-Int"):  # AttentionSplitter, components.py:155-193
        for name, shape in (("input_proj", (128, O_)), ("query_human", (128, 128)), ("query_object", (128, 128)), ("key", (128, 128)),
                            ("value", (128, 128)), ("output_proj", (O_, 128))):
            s[f"attention_splitter.{name}.weight"] = (shape, "w"); s[f"attention_splitter.{name}.bias"] = ((shape[0],), "b")
    return s


def _fan_in(shape, name):
    if name.endswith("output_upscaling.0.weight") or name.endswith("output_upscaling.3.weight"):
        return shape[0]  # ConvTranspose2d k2 s2: every output pixel sums over in-channels only
    n = 1
    for x in shape[1:]:
        n *= x
    return n


# Tensors whose scale is raised so that random-init mask logits have spread (std ~4) instead of ~0
# (SURVEY.md section 0.9 / 8d): with default init every contact probability collapses to 0.5.
_LOGIT_GAIN = {"output_hypernetworks_mlps.0.layers.2.weight": 3.0, "output_upscaling.3.weight": 1.5}


def make_state_dict(cfg, seed: int = 0, device="cpu", gain: float = 1.0):
    """Seeded random-init weights of the architecture, keyed like the reference checkpoint.

    device == 'cpu': numpy PCG64 (identical on every machine), values rounded to bf16-representable fp32 so the
    fp32 oracle and the bf16 product read the same numbers.  On a CUDA device the tensors are drawn on the GPU in
    bf16 (the 13B benchmark weights; no cross-device reproducibility needed there).
    """
    import torch

    spec = state_dict_spec(cfg)
    out = {}
    on_gpu = str(device).startswith("cuda")
    if on_gpu:
        gen = torch.Generator(device=device).manual_seed(seed)
    else:
        rng = np.random.default_rng(seed + 4000)

    def normal(shape, std, mean=0.0):
        if on_gpu:
            t = torch.empty(shape, device=device, dtype=torch.bfloat16)
            t.normal_(mean, std, generator=gen)
            return t
        a = rng.standard_normal(size=shape, dtype=np.float32) * np.float32(std) + np.float32(mean)
        return torch.from_numpy(a).bfloat16().float()

    for name, (shape, kind) in spec.items():
        if kind == "w":
            std = gain / float(np.sqrt(_fan_in(shape, name)))
            for suffix, g in _LOGIT_GAIN.items():
                if name.endswith(suffix):
                    std *= g
            out[name] = normal(shape, std)
        elif kind == "b":
            out[name] = normal(shape, 0.05)
        elif kind == "g":
            out[name] = normal(shape, 0.05, 1.0)
        elif kind == "e":
            out[name] = normal(shape, 0.5 if "embed_tokens" in name else 0.1)
        elif kind == "pe":
            out[name] = normal(shape, 1.0)
        else:
            raise ValueError(kind)
    return out


def make_prompt_ids(cfg, batch: int, n_pre: int = 40, n_post: int = 30, n_answer: int = 24, seed: int = 0):
    """Scripted token protocol.  Returns (input_ids [B,L] with one -200 image token,
    answer_ids [B,G]): prompt = [bos] + n_pre random ids + [<im_start>, -200, <im_end>] + n_post random ids;
    answer = n_answer-3 random ids, [SEG], one random id, </s>."""
    rng = np.random.default_rng(seed + 5000)
    hi = min(cfg.seg_token_idx, cfg.vocab_size)
    ids = np.zeros((batch, 1 + n_pre + 3 + n_post), np.int64)
    ans = np.zeros((batch, n_answer), np.int64)
    for b in range(batch):
        pre = rng.integers(3, hi, n_pre)
        post = rng.integers(3, hi, n_post)
        ids[b] = np.concatenate([[cfg.bos_token_id], pre, [cfg.im_start_token_idx, -200, cfg.im_end_token_idx], post])
        a = rng.integers(3, hi, n_answer)
        a[-3] = cfg.seg_token_idx
        a[-1] = cfg.eos_token_id
        ans[b] = a
    return ids, ans


def make_images(cfg, batch: int, seed: int = 0):
    """uint8 U[0,255] images -> normalised float32: images_clip [B,3,224,224] (CLIP mean/std) and
    images [B,V,3,1024,1024] (SAM mean/std, run_demo.py:65-79)."""
    rng = np.random.default_rng(seed + 6000)
    V, S1, S2 = cfg.multiview_channels, cfg.clip_image_size, cfg.sam_img_size
    clip_u8 = rng.integers(0, 256, size=(batch, 3, S1, S1), dtype=np.uint8)
    sam_u8 = rng.integers(0, 256, size=(batch, V, 3, S2, S2), dtype=np.uint8)
    cm = np.array([0.48145466, 0.4578275, 0.40821073], np.float32).reshape(1, 3, 1, 1)
    cs = np.array([0.26862954, 0.26130258, 0.27577711], np.float32).reshape(1, 3, 1, 1)
    sm = np.array([123.675, 116.28, 103.53], np.float32).reshape(1, 1, 3, 1, 1)
    ss = np.array([58.395, 57.12, 57.375], np.float32).reshape(1, 1, 3, 1, 1)
    clip = (clip_u8.astype(np.float32) / 255.0 - cm) / cs
    sam = (sam_u8.astype(np.float32) - sm) / ss
    return clip, sam
