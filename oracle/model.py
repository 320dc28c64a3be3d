"""Oracle (TEST INFRASTRUCTURE, not product code): CPU restatement of the reference's inference hot path.

Plain torch-on-CPU tensor algebra (F.linear / matmul / softmax), one function per reference function, in the
reference's op order so that running it in bfloat16 rounds where the eager reference rounds and running it in
float32 gives the "exact" answer both implementations approximate.  Weights come as a flat state dict with the
reference's checkpoint key names.

Follows (reference file:line):
  clip_tower            HF CLIPVisionModel via model/llava/model/multimodal_encoder/clip_encoder.py:31-60
                        (hidden_states[select_layer][:, 1:]); transformers 4.31 modeling_clip.py numerics
  encode_images         model/llava/model/llava_arch.py:93-96
  splice_embeddings     model/llava/model/llava_arch.py:98-347 (branch :185-208, one image token per row)
  llama_forward         HF LlamaModel (transformers 4.31 modeling_llama.py) via llava_llama.py:93-105
  seg_embeddings        model/InteractVLM.py:535-576 ([SEG]-1 hidden row -> text_hidden_fcs)
  cam_gate              model/components.py:541-572 + model/InteractVLM.py:268-282
  sam_image_encoder     model/segment_anything/modeling/image_encoder.py:110-426
  dense_pe / prompt     model/segment_anything/modeling/prompt_encoder.py:140-238
  mask_decoder          model/segment_anything/modeling/mask_decoder.py:116-164, transformer.py:62-242
  postprocess_masks     model/segment_anything/modeling/sam.py:137-172
  evaluate              model/InteractVLM.py:510-638 (greedy generate WITHOUT a KV cache)
  model_forward         model/InteractVLM.py:296-474 (inference=True)

Pinned against the UNMODIFIED reference by tests/golden/tiny_model.npz (oracle/make_goldens_model.py) -- see
tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from interactvlm_b200.synthetic import CLIP_PREFIX, SAM_PREFIX

from . import lift as OL

IMAGE_TOKEN_INDEX = -200


class W:
    """State-dict view with a dtype cast (the reference calls model.bfloat16(); fp32 buffers stay fp32)."""

    def __init__(self, sd, dtype=torch.float32, device="cpu"):
        # device="cpu" is the oracle proper; tests may pass a CUDA device to use the same restatement (stock torch
        # fp32 ops) as the checker at full model sizes, where the CPU would take minutes per stage.
        self.sd, self.dtype, self.device = sd, dtype, torch.device(device)
        self._cache = {}

    def __call__(self, name, prefix=""):
        k = prefix + name
        if k not in self._cache:
            self._cache[k] = self.sd[k].detach().to(self.device).to(self.dtype)
        return self._cache[k]

    def has(self, name):
        return name in self.sd


def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


# ---------------------------------------------------------------------------------------------- CLIP
def clip_tower(w: W, cfg, images):
    """images [B,3,224,224] -> [B,256,C]: hidden_states[mm_vision_select_layer] without the CLS row."""
    p = CLIP_PREFIX
    B = images.shape[0]
    x = images.to(w.device, w.dtype)
    pe = F.conv2d(x, w("embeddings.patch_embedding.weight", p), stride=cfg.clip_patch_size)
    pe = pe.flatten(2).transpose(1, 2)
    cls = w("embeddings.class_embedding", p).expand(B, 1, -1)
    h = torch.cat([cls, pe], 1) + w("embeddings.position_embedding.weight", p)[None]
    h = _ln(h, w("pre_layrnorm.weight", p), w("pre_layrnorm.bias", p), cfg.clip_layer_norm_eps)
    nh = cfg.clip_num_attention_heads
    hd = cfg.clip_hidden_size // nh
    for i in range(cfg.clip_layers_used):
        lp = p + f"encoder.layers.{i}."
        r = h
        y = _ln(h, w("layer_norm1.weight", lp), w("layer_norm1.bias", lp), cfg.clip_layer_norm_eps)
        q = F.linear(y, w("self_attn.q_proj.weight", lp), w("self_attn.q_proj.bias", lp)) * (hd ** -0.5)
        k = F.linear(y, w("self_attn.k_proj.weight", lp), w("self_attn.k_proj.bias", lp))
        v = F.linear(y, w("self_attn.v_proj.weight", lp), w("self_attn.v_proj.bias", lp))
        sh = lambda t: t.view(B, -1, nh, hd).transpose(1, 2)
        a = torch.softmax(sh(q) @ sh(k).transpose(-1, -2), -1)
        o = (a @ sh(v)).transpose(1, 2).reshape(B, -1, nh * hd)
        h = r + F.linear(o, w("self_attn.out_proj.weight", lp), w("self_attn.out_proj.bias", lp))
        r = h
        y = _ln(h, w("layer_norm2.weight", lp), w("layer_norm2.bias", lp), cfg.clip_layer_norm_eps)
        y = F.linear(y, w("mlp.fc1.weight", lp), w("mlp.fc1.bias", lp))
        y = y * torch.sigmoid(1.702 * y)  # quick_gelu
        h = r + F.linear(y, w("mlp.fc2.weight", lp), w("mlp.fc2.bias", lp))
    return h[:, 1:]


def encode_images(w: W, cfg, images):
    return F.linear(clip_tower(w, cfg, images), w("model.mm_projector.weight"), w("model.mm_projector.bias"))


# ---------------------------------------------------------------------------------------------- LLaMA
def splice_embeddings(w: W, cfg, input_ids, image_features):
    """ids [B,L] (one -200 each) + image rows [B,256,D] -> [B, L+255, D]."""
    emb = w("model.embed_tokens.weight")
    rows = []
    for b in range(input_ids.shape[0]):
        ids = input_ids[b]
        pos = int((ids == IMAGE_TOKEN_INDEX).nonzero()[0])
        rows.append(torch.cat([emb[ids[:pos]], image_features[b], emb[ids[pos + 1:]]], 0))
    return torch.stack(rows, 0)


def _rope_tables(cfg, S, dtype):
    hd = cfg.head_dim
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    fr = torch.outer(torch.arange(S, dtype=torch.float32), inv)
    emb = torch.cat((fr, fr), -1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), -1)


def _rmsnorm(x, g, eps):
    xf = x.float()
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return g * xf.to(x.dtype)


def llama_forward(w: W, cfg, embeds, trace=None):
    """embeds [B,S,D] -> last hidden state after the final RMSNorm [B,S,D] (causal, no padding).
    trace (optional list): receives the residual stream after every layer (tests: per-layer error curves)."""
    B, S, D = embeds.shape
    nh, hd = cfg.num_attention_heads, cfg.head_dim
    cos, sin = (t.to(embeds.device) for t in _rope_tables(cfg, S, embeds.dtype))
    mask = torch.full((S, S), torch.finfo(embeds.dtype).min, dtype=embeds.dtype, device=embeds.device).triu(1)
    h = embeds
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}."
        r = h
        y = _rmsnorm(h, w("input_layernorm.weight", p), cfg.rms_norm_eps)
        sh = lambda t: t.view(B, S, nh, hd).transpose(1, 2)
        q = sh(F.linear(y, w("self_attn.q_proj.weight", p)))
        k = sh(F.linear(y, w("self_attn.k_proj.weight", p)))
        v = sh(F.linear(y, w("self_attn.v_proj.weight", p)))
        q = q * cos + _rotate_half(q) * sin
        k = k * cos + _rotate_half(k) * sin
        a = (q @ k.transpose(-1, -2)) / math.sqrt(hd) + mask
        a = torch.softmax(a, -1, dtype=torch.float32).to(q.dtype)
        o = (a @ v).transpose(1, 2).reshape(B, S, D)
        h = r + F.linear(o, w("self_attn.o_proj.weight", p))
        r = h
        y = _rmsnorm(h, w("post_attention_layernorm.weight", p), cfg.rms_norm_eps)
        y = F.silu(F.linear(y, w("mlp.gate_proj.weight", p))) * F.linear(y, w("mlp.up_proj.weight", p))
        h = r + F.linear(y, w("mlp.down_proj.weight", p))
        if trace is not None:
            trace.append(h)
    return _rmsnorm(h, w("model.norm.weight"), cfg.rms_norm_eps)


def lm_logits(w: W, hidden):
    return F.linear(hidden, w("lm_head.weight"))


def seg_token_ids(cfg):
    """[SEG] alone, or [SEG] / [HSEG] / [OSEG] for the Gen-Hu-Obj / Gen-Int token types (InteractVLM.py:535-543)."""
    ids = [cfg.seg_token_idx]
    if cfg.token_type.replace("-DifDe", "") in ("Gen-Hu-Obj", "Gen-Int"):
        ids += [t for t in (cfg.hseg_token_idx, cfg.oseg_token_idx) if t is not None]
    return ids


def seg_rows(cfg, output_ids, with_tokens=False):
    """Hidden-state row that predicts each segmentation token: token j sits at row j + img_emb_len, its predictor at
    j - 1 + img_emb_len (InteractVLM.py:535-549).  Returns a list (per sample) of row indices; with_tokens also the
    token id the reference files per sample (the first one, :566-575)."""
    out, toks = [], []
    ids = torch.tensor(seg_token_ids(cfg))
    for b in range(output_ids.shape[0]):
        js = [j for j in torch.isin(output_ids[b].cpu(), ids).nonzero().flatten().tolist() if j >= 1]
        out.append([j - 1 + cfg.img_emb_len for j in js])
        toks.append(int(output_ids[b, js[0]]) if js else None)
    return (out, toks) if with_tokens else out


def text_hidden_fcs(w: W, x):
    y = F.relu(F.linear(x, w("model.text_hidden_fcs.0.0.weight"), w("model.text_hidden_fcs.0.0.bias")))
    return F.linear(y, w("model.text_hidden_fcs.0.2.weight"), w("model.text_hidden_fcs.0.2.bias"))


def cam_gate(w: W, cfg, pred_emb, cam_params):
    """pred_emb [N,256], cam_params [V,5] -> [N,V,256]: the camera conditioning of process_embeddings
    (InteractVLM.py:268-283) for the three encoder types of components.py:491-572 --
    simple: emb + relu(W cam); view_index: emb * W_v sigmoid(W2 relu(W1 cam)); vi_v1: emb * sigmoid(W_v relu(W2 relu(W1 cam)))."""
    V = cfg.multiview_channels
    emb = pred_emb[:, None, :].repeat(1, V, 1)
    if not cfg.multiview_cam_cond:
        return emb
    lin = lambda x, name: F.linear(x, w(f"cam_pose_encoder.{name}.weight"), w(f"cam_pose_encoder.{name}.bias"))
    cam = cam_params.to(w.device, w.dtype)
    if cfg.cam_encoder_type == "simple":
        return emb + F.relu(lin(cam, "linear1"))
    assert cfg.cam_encoder_type in ("view_index", "vi_v1"), cfg.cam_encoder_type
    encs = []
    for v in range(V):
        c = cam[[v]]
        if cfg.cam_encoder_type == "view_index":
            y = torch.sigmoid(lin(F.relu(lin(c, "spatial_encoder.0")), "spatial_encoder.2"))
            y = lin(y, f"view_transforms.{v}")
        else:
            y = F.relu(lin(F.relu(lin(c, "spatial_encoder.0")), "spatial_encoder.2"))
            y = torch.sigmoid(lin(y, f"view_transforms.{v}"))
        encs.append(y)
    return emb * torch.stack(encs, 1)


def attention_split(w: W, x, which):
    """AttentionSplitter.forward (components.py:173-193) on x [N,V,256]: the V view tokens attend to each other with a
    human or an object query projection.  which: 'human' | 'object'."""
    lin = lambda t, name: F.linear(t, w(f"attention_splitter.{name}.weight"), w(f"attention_splitter.{name}.bias"))
    xp = lin(x, "input_proj")
    k, v = lin(xp, "key"), lin(xp, "value")
    q = lin(xp, "query_human" if which == "human" else "query_object")
    attn = F.softmax(torch.matmul(q, k.transpose(-2, -1)) / (k.size(-1) ** 0.5), dim=-1)
    return lin(torch.matmul(attn, v), "output_proj")


def process_embeddings(w: W, cfg, pred_emb, cam_params, token):
    """InteractVLM.py:268-294: camera conditioning, then (Gen-Hu-Obj / Gen-Int only) the splitter branch the token selects."""
    e = cam_gate(w, cfg, pred_emb, cam_params)
    if cfg.token_type.replace("-DifDe", "") == "Gen":
        return e
    if token is not None and token == cfg.hseg_token_idx:
        return attention_split(w, e, "human")
    if token is not None and token == cfg.oseg_token_idx:
        return attention_split(w, e, "object")
    return e


# ---------------------------------------------------------------------------------------------- SAM encoder
def _rel_pos(size, rel_pos):
    idx = torch.arange(size)[:, None] - torch.arange(size)[None, :] + (size - 1)
    return rel_pos[idx.to(rel_pos.device)]  # [q, k, hd]; table length is always 2*size-1 on this path (no interpolation)


def _sam_attention(w: W, p, x, nh):
    B, H, Wd, E = x.shape
    hd = E // nh
    qkv = F.linear(x, w("attn.qkv.weight", p), w("attn.qkv.bias", p)).reshape(B, H * Wd, 3, nh, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.reshape(3, B * nh, H * Wd, hd).unbind(0)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    Rh, Rw = _rel_pos(H, w("attn.rel_pos_h", p)), _rel_pos(Wd, w("attn.rel_pos_w", p))
    rq = q.reshape(B * nh, H, Wd, hd)
    rel_h = torch.einsum("bhwc,hkc->bhwk", rq, Rh)
    rel_w = torch.einsum("bhwc,wkc->bhwk", rq, Rw)
    attn = (attn.view(B * nh, H, Wd, H, Wd) + rel_h[:, :, :, :, None] + rel_w[:, :, :, None, :]).view(B * nh, H * Wd, H * Wd)
    attn = attn.softmax(-1)
    o = (attn @ v).view(B, nh, H, Wd, hd).permute(0, 2, 3, 1, 4).reshape(B, H, Wd, E)
    return F.linear(o, w("attn.proj.weight", p), w("attn.proj.bias", p))


def _ln2d(x, g, b, eps=1e-6):
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return g[:, None, None] * x + b[:, None, None]


def sam_image_encoder(w: W, cfg, images, trace=None):
    """images [N,3,1024,1024] -> [N,256,64,64].  trace (optional list): the token grid after every block."""
    e = SAM_PREFIX + "image_encoder."
    x = F.conv2d(images.to(w.device, w.dtype), w("patch_embed.proj.weight", e), w("patch_embed.proj.bias", e),
                 stride=cfg.sam_patch_size).permute(0, 2, 3, 1)
    x = x + w("pos_embed", e)
    ws = cfg.sam_window_size
    for i in range(cfg.sam_depth):
        p = e + f"blocks.{i}."
        shortcut = x
        y = _ln(x, w("norm1.weight", p), w("norm1.bias", p), 1e-6)
        if i not in cfg.sam_global_attn_indexes:
            B, H, Wd, E = y.shape
            ph, pw = (ws - H % ws) % ws, (ws - Wd % ws) % ws
            y = F.pad(y, (0, 0, 0, pw, 0, ph))
            Hp, Wp = H + ph, Wd + pw
            y = y.view(B, Hp // ws, ws, Wp // ws, ws, E).permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, E)
            y = _sam_attention(w, p, y, cfg.sam_num_heads)
            y = y.view(B, Hp // ws, Wp // ws, ws, ws, E).permute(0, 1, 3, 2, 4, 5).contiguous().view(B, Hp, Wp, E)
            y = y[:, :H, :Wd, :].contiguous()
        else:
            y = _sam_attention(w, p, y, cfg.sam_num_heads)
        x = shortcut + y
        y = _ln(x, w("norm2.weight", p), w("norm2.bias", p), 1e-6)
        y = F.linear(F.gelu(F.linear(y, w("mlp.lin1.weight", p), w("mlp.lin1.bias", p))), w("mlp.lin2.weight", p),
                     w("mlp.lin2.bias", p))
        x = x + y
        if trace is not None:
            trace.append(x)
    x = x.permute(0, 3, 1, 2)
    x = F.conv2d(x, w("neck.0.weight", e))
    x = _ln2d(x, w("neck.1.weight", e), w("neck.1.bias", e))
    x = F.conv2d(x, w("neck.2.weight", e), padding=1)
    return _ln2d(x, w("neck.3.weight", e), w("neck.3.bias", e))


# ---------------------------------------------------------------------------------------------- prompt + decoder
def dense_pe(w: W, cfg):
    """PositionEmbeddingRandom.forward((64,64)) -> [1,256,64,64]; the gaussian matrix is a buffer that
    model.bfloat16() casts too, so the whole encoding runs in the model dtype (prompt_encoder.py:203-229)."""
    g = cfg.sam_grid
    G = w(SAM_PREFIX + "prompt_encoder.pe_layer.positional_encoding_gaussian_matrix")
    grid = torch.ones((g, g), dtype=G.dtype, device=G.device)
    y = (grid.cumsum(0) - 0.5) / g
    x = (grid.cumsum(1) - 0.5) / g
    c = 2 * torch.stack([x, y], -1) - 1
    c = c @ G
    c = 2 * np.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], -1).permute(2, 0, 1)[None]


def _dec_attn(w: W, p, q, k, v, nh):
    q = F.linear(q, w("q_proj.weight", p), w("q_proj.bias", p))
    k = F.linear(k, w("k_proj.weight", p), w("k_proj.bias", p))
    v = F.linear(v, w("v_proj.weight", p), w("v_proj.bias", p))
    sep = lambda t: t.reshape(t.shape[0], t.shape[1], nh, t.shape[2] // nh).transpose(1, 2)
    q, k, v = sep(q), sep(k), sep(v)
    a = q @ k.permute(0, 1, 3, 2)
    a = a / math.sqrt(q.shape[-1])
    a = torch.softmax(a, -1)
    o = a @ v
    o = o.transpose(1, 2).reshape(o.shape[0], o.shape[2], -1)
    return F.linear(o, w("out_proj.weight", p), w("out_proj.bias", p))


def decoder_name(cfg, ds_name):
    """ModifiedSAM.forward (InteractVLM.py:44-53): with token_type '*-DifDe' the human / object copies of the mask decoder
    (initialize_separate_decoders, :114-122) serve 'hcontact' / 'oafford' + 'ocontact'; everything else the shared one."""
    if "DifDe" in cfg.token_type and ds_name is not None:
        if "hcontact" in ds_name:
            return "human_mask_decoder"
        if "oafford" in ds_name or "ocontact" in ds_name:
            return "object_mask_decoder"
    return "mask_decoder"


def mask_decoder(w: W, cfg, image_embeddings, sparse, which="mask_decoder"):
    """image_embeddings [V,256,64,64], sparse [1,V,256] -> low-res logits [V,1,256,256] (multimask_output=False)."""
    d = SAM_PREFIX + which + "."
    nh = cfg.sam_dec_heads
    nm = cfg.sam_num_multimask_outputs + 1
    out_tok = torch.cat([w("iou_token.weight", d), w("mask_tokens.weight", d)], 0)[None].expand(sparse.shape[0], -1, -1)
    tokens = torch.cat((out_tok, sparse), 1)
    src = torch.repeat_interleave(image_embeddings, tokens.shape[0], 0)
    dense = w(SAM_PREFIX + "prompt_encoder.no_mask_embed.weight").reshape(1, -1, 1, 1)
    src = src + dense
    pos = torch.repeat_interleave(dense_pe(w, cfg), tokens.shape[0], 0)
    b, c, h, wd = src.shape
    keys = src.flatten(2).permute(0, 2, 1)
    key_pe = pos.flatten(2).permute(0, 2, 1)
    queries, query_pe = tokens, tokens
    for i in range(cfg.sam_dec_depth):
        p = d + f"transformer.layers.{i}."
        if i == 0:
            queries = _dec_attn(w, p + "self_attn.", queries, queries, queries, nh)
        else:
            q = queries + query_pe
            queries = queries + _dec_attn(w, p + "self_attn.", q, q, queries, nh)
        queries = _ln(queries, w("norm1.weight", p), w("norm1.bias", p), 1e-5)
        q, k = queries + query_pe, keys + key_pe
        queries = queries + _dec_attn(w, p + "cross_attn_token_to_image.", q, k, keys, nh)
        queries = _ln(queries, w("norm2.weight", p), w("norm2.bias", p), 1e-5)
        m = F.linear(F.relu(F.linear(queries, w("mlp.lin1.weight", p), w("mlp.lin1.bias", p))), w("mlp.lin2.weight", p),
                     w("mlp.lin2.bias", p))
        queries = _ln(queries + m, w("norm3.weight", p), w("norm3.bias", p), 1e-5)
        q, k = queries + query_pe, keys + key_pe
        keys = keys + _dec_attn(w, p + "cross_attn_image_to_token.", k, q, queries, nh)
        keys = _ln(keys, w("norm4.weight", p), w("norm4.bias", p), 1e-5)
    q, k = queries + query_pe, keys + key_pe
    queries = queries + _dec_attn(w, d + "transformer.final_attn_token_to_image.", q, k, keys, nh)
    hs = _ln(queries, w("transformer.norm_final_attn.weight", d), w("transformer.norm_final_attn.bias", d), 1e-5)
    mask_tokens_out = hs[:, 1:1 + nm, :]
    src = keys.transpose(1, 2).view(b, c, h, wd)
    up = F.conv_transpose2d(src, w("output_upscaling.0.weight", d), w("output_upscaling.0.bias", d), stride=2)
    up = F.gelu(_ln2d(up, w("output_upscaling.1.weight", d), w("output_upscaling.1.bias", d)))
    up = F.gelu(F.conv_transpose2d(up, w("output_upscaling.3.weight", d), w("output_upscaling.3.bias", d), stride=2))
    hyper = []
    for i in range(nm):
        x = mask_tokens_out[:, i, :]
        hp = d + f"output_hypernetworks_mlps.{i}.layers."
        x = F.relu(F.linear(x, w("0.weight", hp), w("0.bias", hp)))
        x = F.relu(F.linear(x, w("1.weight", hp), w("1.bias", hp)))
        hyper.append(F.linear(x, w("2.weight", hp), w("2.bias", hp)))
    hyper = torch.stack(hyper, 1)
    b, c, h, wd = up.shape
    masks = (hyper @ up.view(b, c, h * wd)).view(b, nm, h, wd)
    return masks[:, 0:1]


def postprocess_masks(cfg, masks, input_size, original_size):
    m = F.interpolate(masks.float(), (cfg.sam_img_size, cfg.sam_img_size), mode="bilinear", align_corners=False)
    m = m[..., : input_size[0], : input_size[1]]
    return F.interpolate(m, tuple(original_size), mode="bilinear", align_corners=False)


# ---------------------------------------------------------------------------------------------- whole path
def lm_hidden(w: W, cfg, images_clip, ids):
    feats = encode_images(w, cfg, images_clip)
    return llama_forward(w, cfg, splice_embeddings(w, cfg, ids, feats))


def greedy_generate(w: W, cfg, images_clip, input_ids, max_new_tokens, scripted=None):
    """HF greedy search as the reference runs it: no KV cache, the full sequence is re-encoded every step
    (SURVEY.md section 0.3).  `scripted` [B,G] forces the next tokens (teacher forcing) for arithmetic parity.
    Returns (sequences [B,L+g], last-step hidden states [B, L+g-1+255, D], per-step greedy tokens)."""
    ids = input_ids.clone()
    B = ids.shape[0]
    done = torch.zeros(B, dtype=torch.bool)
    greedy = []
    hidden = None
    for step in range(max_new_tokens):
        hidden = lm_hidden(w, cfg, images_clip, ids)
        nxt = lm_logits(w, hidden[:, -1]).float().argmax(-1)
        greedy.append(nxt.clone())
        if scripted is not None:
            nxt = scripted[:, step].clone()
        nxt = torch.where(done, torch.full_like(nxt, cfg.pad_token_id), nxt)
        ids = torch.cat([ids, nxt[:, None]], 1)
        done |= nxt == cfg.eos_token_id
        if bool(done.all()):
            break
    return ids, hidden, torch.stack(greedy, 1)


def masks_from_hidden(w: W, cfg, hidden, output_ids, images, cam_params, resize_list, original_size_list, stages=None,
                      ds_names=None):
    """Everything downstream of the language model (InteractVLM.py:535-612): per sample a [V,H,W] fp32 logit map.
    ds_names: per-sample dataset / contact-type names (ds_name_list[i] at :435, contact_type at :604) -- they select the
    decoder copy of the '*-DifDe' token types."""
    rows, tokens = seg_rows(cfg, output_ids, with_tokens=True)
    pred_masks = []
    for b in range(hidden.shape[0]):
        emb_img = sam_image_encoder(w, cfg, images[b])
        pe = text_hidden_fcs(w, hidden[b, rows[b]]) if rows[b] else hidden.new_zeros((0, cfg.out_dim))
        if stages is not None:
            stages.setdefault("sam_embeddings", []).append(emb_img)
            stages.setdefault("pred_embeddings", []).append(pe)
        if pe.shape[0] == 0:
            pred_masks.append(torch.zeros((0,) + tuple(original_size_list[b]), dtype=torch.float32))
            continue
        assert pe.shape[0] == 1, "multi-view decoding broadcasts only for one [SEG] per sample (SURVEY.md 0.5)"
        prompt = process_embeddings(w, cfg, pe, cam_params[b], tokens[b])
        low = mask_decoder(w, cfg, emb_img, prompt, decoder_name(cfg, ds_names[b] if ds_names is not None else None))
        pm = postprocess_masks(cfg, low, resize_list[b], original_size_list[b])
        if stages is not None:
            stages.setdefault("prompt", []).append(prompt)
            stages.setdefault("low_res", []).append(low)
        pred_masks.append(pm[:, 0])
    return pred_masks


def evaluate(sd, cfg, images_clip, images, input_ids, cam_params, resize_list, original_size_list, lift_maps=None,
             contact_type="hcontact", max_new_tokens=32, scripted=None, dtype=torch.float32, stages=None):
    """InteractVLMForCausalLM.evaluate (InteractVLM.py:510-638).  lift_maps = (p2v [V,H,W,3], bary [V,H,W,3], n_verts)."""
    w = W(sd, dtype)
    with torch.no_grad():
        output_ids, hidden, greedy = greedy_generate(w, cfg, images_clip.to(dtype), input_ids, max_new_tokens, scripted)
        if stages is not None:
            stages["hidden"] = hidden
            stages["greedy"] = greedy
        pred_masks = masks_from_hidden(w, cfg, hidden, output_ids, images.to(dtype), cam_params.to(dtype), resize_list,
                                       original_size_list, stages, ds_names=[contact_type] * output_ids.shape[0])
    contact = None
    if pred_masks[0].shape[0] > 0 and lift_maps is not None:
        p2v, bary, n = lift_maps
        m = np.stack([pm.numpy() for pm in pred_masks], 0)
        if "hcontact" in contact_type and cfg.hC_loss_weight > 0:
            contact = torch.from_numpy(OL.lift_human(m, p2v, bary, n))
        else:
            contact = torch.from_numpy(OL.lift_object_mesh(m, p2v, bary, n, thr=0.3))
    return {"output_ids": output_ids, "pred_masks": pred_masks, "pred_contact_3d": contact}


def model_forward(sd, cfg, images, images_clip, input_ids, cam_params, resize_list, label_shapes, lift_maps=None,
                  dtype=torch.float32, stages=None, ds_name_list=None):
    """model_forward(inference=True) (InteractVLM.py:296-474): one teacher-forced pass over prompt+answer."""
    w = W(sd, dtype)
    with torch.no_grad():
        hidden = lm_hidden(w, cfg, images_clip.to(dtype).expand(input_ids.shape[0], -1, -1, -1), input_ids)
        if stages is not None:
            stages["hidden"] = hidden
        pred_masks = masks_from_hidden(w, cfg, hidden, input_ids, images.to(dtype), cam_params.to(dtype), resize_list,
                                       label_shapes, stages, ds_names=ds_name_list)
    out = {"pred_masks": pred_masks}
    if lift_maps is not None and cfg.hC_loss_weight > 0:
        p2v, bary, n = lift_maps
        out["pred_human_3d_contact"] = torch.from_numpy(
            OL.lift_human(np.stack([pm.numpy() for pm in pred_masks], 0), p2v, bary, n))
    return out
