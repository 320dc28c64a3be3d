"""Caller-side pieces of the reference's harnesses that sit directly on either side of the hot path:

  * input preparation of `run_demo.py:330-379` with the normalise / pad / cast fused on the GPU (`ops.preprocess_u8`);
  * `convert_contacts` (utils/utils.py:428-443) as a CSR SpMV, and the on-disk result formats of `run_demo.py:438-463`
    (`*_hcontact_vertices.npz` with `pred_contact_3d_smplh` / `pred_contact_3d_smplx`, `*_oafford_vertices.npz` with
    `pred_contact_3d`);
  * a batched `validate()` in the shape of `evaluate.py:41-302`: per-rank shard of the samples, `model.evaluate()` on
    batches (the reference is batch 1), `get_h_contact_metrics` (utils/eval_utils.py:63-94) and ONE all-gather of the
    predictions instead of all_gather(list) + all_gather_object + per-meter all_reduce (evaluate.py:185-222).

Dataset classes, tokenisation and image decoding stay with the caller (they need the downloaded datasets).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from .parallel import gather_contacts, shard_range


def prepare_inputs(model, image_u8, sam_views_u8):
    """uint8 HWC images (already resized like the reference: CLIP 224x224 crop, SAM longest side 1024) ->
    (images_clip [B,3,224,224] bf16, images [B,V,3,1024,1024] bf16, resize_list).
    image_u8 [B,224,224,3]; sam_views_u8 [B,V,h,w,3] with h,w <= 1024 (zero padded to 1024 like `preprocess`)."""
    cfg, ctx, dev = model.config, model.ctx, model.device
    image_u8 = torch.as_tensor(image_u8).to(dev).contiguous()
    sam_views_u8 = torch.as_tensor(sam_views_u8).to(dev).contiguous()
    B, V, h, w, _ = sam_views_u8.shape
    clip = ctx.preprocess_u8(image_u8, cfg.clip_image_size, kind="clip")
    sam = ctx.preprocess_u8(sam_views_u8.view(B * V, h, w, 3), cfg.sam_img_size, kind="sam")
    return clip, sam.view(B, V, 3, cfg.sam_img_size, cfg.sam_img_size), [(h, w)] * B


def prepare_inputs_from_raw(model, image_u8, sam_views_u8):
    """Same as prepare_inputs but starting from UN-resized uint8 images, with the reference's CPU/Pillow resizes done on the
    GPU bit-exactly: CLIP = bicubic to shortest edge 224 + centre crop (CLIPImageProcessor, run_demo.py:330-346), SAM =
    bilinear to longest side 1024 (ResizeLongestSide.apply_image, run_demo.py:359-366).
    image_u8 [B,H,W,3]; sam_views_u8 [B,V,h,w,3] (all views of a call share one size)."""
    from . import resample as R

    cfg, ctx, dev = model.config, model.ctx, model.device
    image_u8 = torch.as_tensor(image_u8).to(dev).contiguous()
    sam_views_u8 = torch.as_tensor(sam_views_u8).to(dev).contiguous()
    B, H, W, _ = image_u8.shape
    ch, cw = R.clip_target_size(H, W, cfg.clip_image_size)
    small = ctx.resize_u8(image_u8, ch, cw, "bicubic")
    top, left = (ch - cfg.clip_image_size) // 2, (cw - cfg.clip_image_size) // 2
    crop = small[:, top:top + cfg.clip_image_size, left:left + cfg.clip_image_size].contiguous()
    clip = ctx.preprocess_u8(crop, cfg.clip_image_size, kind="clip")
    _, V, h, w, _ = sam_views_u8.shape
    sh, sw = R.sam_target_size(h, w, cfg.sam_img_size)
    views = ctx.resize_u8(sam_views_u8.view(B * V, h, w, 3), sh, sw, "bilinear")
    sam = ctx.preprocess_u8(views, cfg.sam_img_size, kind="sam")
    return clip, sam.view(B, V, 3, cfg.sam_img_size, cfg.sam_img_size), [(sh, sw)] * B


def prepare_inputs_from_jpeg(model, image_jpeg: bytes, sam_views_u8):
    """run_demo.py:330-366 starting from the COMPRESSED photo: nvJPEG decode on the GPU (ivlm_jpeg_decode_rgb), then the
    Pillow-exact resizes and the fused normalisation of prepare_inputs_from_raw.  One photo, V rendered views.  The decoded
    pixels can differ from cv2.imread's by a few grey levels (different IDCT); everything after the decode is bit-exact."""
    rgb = model.ctx.decode_jpeg(image_jpeg)
    return prepare_inputs_from_raw(model, rgb[None], torch.as_tensor(sam_views_u8)[None] if torch.as_tensor(sam_views_u8).dim() == 4
                                   else sam_views_u8)


class ContactConverter:
    """convert_contacts(contact, mapping): dense [n_out, n_in] mapping applied as a CSR SpMV (the SMPL->SMPL-X matrix has
    ~3 non-zeros per row; the reference streams the 289 MB dense matrix through bmm on every call)."""

    def __init__(self, model, mapping: np.ndarray):
        self.model = model
        if getattr(model, "_emulated", False):
            self._dense = torch.as_tensor(np.asarray(mapping, dtype=np.float32))
            self._csr = None
        else:
            from .ops import CsrMatrix

            self._csr = CsrMatrix(model.ctx, np.asarray(mapping, dtype=np.float32))

    def __call__(self, contact: torch.Tensor) -> torch.Tensor:
        x = contact.float().contiguous()
        y = self._csr(x) if self._csr is not None else (self._dense @ x.T).T
        return y.squeeze()  # the reference's .squeeze(): [10475] for a single sample


def save_hcontact(path_prefix, pred_contact_3d, pred_contact_3d_smplx):
    """`{dir}/{fname}_hcontact_vertices.npz` (run_demo.py:449-452)."""
    out = Path(f"{path_prefix}_hcontact_vertices.npz")
    np.savez(out, pred_contact_3d_smplh=pred_contact_3d.detach().cpu().numpy(),
             pred_contact_3d_smplx=pred_contact_3d_smplx.detach().cpu().numpy())
    return out


def save_ocontact(path_prefix, pred_contact_3d):
    """`{dir}/{fname}_oafford_vertices.npz` (run_demo.py:463)."""
    out = Path(f"{path_prefix}_oafford_vertices.npz")
    np.savez(out, pred_contact_3d=pred_contact_3d.detach().cpu().numpy())
    return out


def h_contact_metrics(contact_gt: torch.Tensor, contact_pred: torch.Tensor, threshold: float = 0.5):
    """get_h_contact_metrics (utils/eval_utils.py:63-94): batch means of F1 / precision / recall."""
    gt = (contact_gt.float() > 0).float()
    pr = (contact_pred.float() >= threshold).float()
    tp = (pr * gt).sum(-1)
    precision = tp / (pr.sum(-1) + 1e-10)
    recall = tp / (gt.sum(-1) + 1e-10)
    f1 = 2 * precision * recall / (precision + recall + 1e-10)
    return f1.mean().item(), precision.mean().item(), recall.mean().item()


def validate(model, samples, batch_size=8, dist=None, max_new_tokens=32, contact_type="hcontact"):
    """Batched evaluation over `samples` (a sequence of dicts with the keys evaluate() takes: images_clip, images,
    input_ids, cam_params, resize, original_size, and optionally gt_contact_3d / scripted).  Samples are sharded
    contiguously over the ranks of `dist` (torch.distributed or None); every rank returns the predictions of ALL samples
    (one all-gather) and the metrics over those that carry a ground truth.  Prompts inside one batch must have equal
    length (batches are formed from consecutive samples of equal prompt length)."""
    n = len(samples)
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    lo, hi = shard_range(n, rank, world)
    preds = []
    hmap = getattr(model.human_3d_contact_predictor, "map", None)
    n_verts = hmap.n if hmap is not None else 6890
    i = lo
    while i < hi:
        j = min(i + batch_size, hi)
        chunk = samples[i:j]
        cat = lambda k: torch.stack([torch.as_tensor(s[k]) for s in chunk], 0)
        scripted = cat("scripted") if all("scripted" in s for s in chunk) else None
        # prompts of different lengths (object names differ) go in as a list: generate() right-pads them like the reference's
        # collate_fn (datasets/dataset.py:159-178) and keeps per-sample positions
        ids = [torch.as_tensor(s["input_ids"]).reshape(-1) for s in chunk]
        out = model.evaluate(cat("images_clip"), cat("images"), ids, cat("cam_params"),
                             [tuple(s["resize"]) for s in chunk], [tuple(s["original_size"]) for s in chunk],
                             contact_type=contact_type, max_new_tokens=max_new_tokens, scripted=scripted)
        pc = out["pred_contact_3d"]
        if pc is None:   # no sample of the chunk produced a [SEG] (the reference reads .get("pred_contact_3d", None), evaluate.py:111)
            pc = torch.zeros((len(chunk), n_verts), device=model.device, dtype=torch.float32)
        preds.append(pc)
        i = j
    n_out = preds[0].shape[1] if preds else n_verts
    local = torch.cat(preds, 0) if preds else torch.zeros((0, n_out), device=model.device)
    allp = gather_contacts(local, dist, n_samples=n) if world > 1 else local
    gts = [s.get("gt_contact_3d") for s in samples]
    metrics = None
    if all(g is not None for g in gts) and n > 0:
        gt = torch.stack([torch.as_tensor(g) for g in gts], 0).to(allp.device)
        f1, p, r = h_contact_metrics(gt, allp)
        metrics = {"f1": f1, "precision": p, "recall": r, "n": n}
    return allp, metrics


# ------------------------------------------------------------------------------------------------ prompt / camera helpers
# Host-side input preparation of both harnesses: plain Python, no device work.
IMAGE_TOKEN_INDEX = -200                       # utils/utils.py:18-23
DEFAULT_IMAGE_TOKEN = "<image>"
DEFAULT_IM_START_TOKEN = "<im_start>"
DEFAULT_IM_END_TOKEN = "<im_end>"
HCONTACT_PROMPT = "Which body parts are in contact with the {object}? Segment these contact areas."   # run_demo.py:282

# preprocess_data/constants.py:315-382: (distance, elevation, azimuth, x_translation, y_translation) per body render
HUMAN_VIEW_CAM_PARAMS = {
    view_type: {"topfront": [2., 45., 315., 0., 0.], "bottomfront": [2., 315., 315., 0., 0.3],
                "topback": [2., 45., 135., 0., 0.], "bottomback": [2., 315., 135., 0., 0.3]}
    for view_type in ("4MV-Z_Vitru", "4MV-Z_Vitru_mv2", "4MV-Z_Vitru_FootGround")}

_SYSTEM = {
    "llava_v1": "A chat between a curious human and an artificial intelligence assistant. "
                "The assistant gives helpful, detailed, and polite answers to the human's questions.",
    "llava_llama_2": "You are a helpful language and vision assistant. You are able to understand the visual content that "
                     "the user provides, and assist the user with a variety of tasks using natural language.",
}


def normalize_cam_params(cam_params):
    """datasets/base_contact_dataset.py:37-50 -> float32 tensor [5]; None -> zeros."""
    if cam_params is None:
        return torch.tensor([0.0, 0.0, 0.0, 0.0, 0.0])
    distance, elevation, azimuth, x_translation, y_translation = cam_params
    return torch.tensor([distance / 10.0, elevation / 360.0, azimuth / 360.0, (x_translation + 1.0) / 2.0,
                         (y_translation + 1.0) / 2.0])


def human_cam_params(view_type="4MV-Z_Vitru"):
    """run_demo.py:275-278: the [1,V,5] camera conditioning of the hcontact path."""
    views = HUMAN_VIEW_CAM_PARAMS[view_type]
    return torch.stack([normalize_cam_params(views[v]) for v in views]).unsqueeze(0)


def build_prompt(question, conv_type="llava_v1", use_mm_start_end=True):
    """run_demo.py:313-323: `<image>\\n` + question, image token wrapped in <im_start>/<im_end>, one user turn and an empty
    assistant turn in the `llava_v1` (SeparatorStyle.TWO) or `llava_llama_2` template (model/llava/conversation.py)."""
    prompt = DEFAULT_IMAGE_TOKEN + "\n" + question
    if use_mm_start_end:
        prompt = prompt.replace(DEFAULT_IMAGE_TOKEN, DEFAULT_IM_START_TOKEN + DEFAULT_IMAGE_TOKEN + DEFAULT_IM_END_TOKEN)
    if conv_type == "llava_v1":
        return _SYSTEM[conv_type] + " " + "USER: " + prompt + " " + "ASSISTANT:"
    if conv_type == "llava_llama_2":
        return "[INST] " + f"<<SYS>>\n{_SYSTEM[conv_type]}\n<</SYS>>\n\n" + prompt + " [/INST]"
    raise ValueError(f"unknown conv_type {conv_type!r} (run_demo.py accepts llava_v1 and llava_llama_2)")


def tokenizer_image_token(prompt, tokenizer, image_token_index=IMAGE_TOKEN_INDEX, return_tensors=None):
    """model/llava/mm_utils.py:19-44: tokenise the text around every `<image>` and put `image_token_index` in between; a
    leading BOS is kept once."""
    chunks = [tokenizer(chunk).input_ids for chunk in prompt.split(DEFAULT_IMAGE_TOKEN)]
    has_bos = len(chunks) > 0 and len(chunks[0]) > 0 and chunks[0][0] == tokenizer.bos_token_id
    input_ids = [chunks[0][0]] if has_bos else []
    skip = 1 if has_bos else 0
    for k, chunk in enumerate(chunks):
        if k > 0:
            input_ids.append(image_token_index)   # the reference inserts [idx] * (offset + 1) and drops `offset` of them
        input_ids.extend(chunk[skip:])
    if return_tensors is not None:
        if return_tensors == "pt":
            return torch.tensor(input_ids, dtype=torch.long)
        raise ValueError(f"Unsupported tensor type: {return_tensors}")
    return input_ids
