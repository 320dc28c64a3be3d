"""Golden vectors for the model stages, produced by the UNMODIFIED reference imported from /root/reference under
oracle/ref_shim.py (test infrastructure; runs only in the build container -- the GPU box has no reference).

    python -m oracle.make_goldens tiny

builds the reference InteractVLMForCausalLM at IVLMConfig.tiny() sizes (SAM via the reference's own _build_sam),
loads the seeded synthetic state dict (interactvlm_b200/synthetic.py), and records
  * model_forward(inference=True) on prompt + scripted answer        (InteractVLM.py:296-474)
  * evaluate() through a no-KV-cache greedy/scripted generate shim    (InteractVLM.py:510-638)
in float32 and in bfloat16 (model.bfloat16(), as evaluate.py:532 mandates).  Large outputs are sub-sampled.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"

from interactvlm_b200 import synthetic as S  # noqa: E402
from interactvlm_b200.config import IVLMConfig  # noqa: E402
from oracle import ref_shim  # noqa: E402

TINY_SEED = dict(weights=0, maps=0, images=11, prompt=12)
LOW_STRIDE, FULL_STRIDE, EMB_STRIDE = 4, 16, 8


def tiny_inputs(cfg, batch=1):
    ids, ans = S.make_prompt_ids(cfg, batch, n_pre=10, n_post=8, n_answer=6, seed=TINY_SEED["prompt"])
    clip, sam = S.make_images(cfg, batch, seed=TINY_SEED["images"])
    cam = np.broadcast_to(S.HCONTACT_CAM_PARAMS, (batch,) + S.HCONTACT_CAM_PARAMS.shape).copy()
    bf = lambda a: torch.from_numpy(a).bfloat16().float()  # bf16-representable inputs for both sides
    return torch.from_numpy(ids), torch.from_numpy(ans), bf(clip), bf(sam), bf(cam)


def build_reference(cfg: IVLMConfig, sd):
    """The reference model object at `cfg` sizes with `sd` loaded (fp32, eval)."""
    p2v, bary = S.make_mesh_lift_maps(seed=TINY_SEED["maps"])
    ref_shim.write_human_maps(p2v, bary)
    clipcfg = ref_shim.tiny_clip_config(hidden=cfg.clip_hidden_size, layers=cfg.clip_num_hidden_layers,
                                        heads=cfg.clip_num_attention_heads, mlp=cfg.clip_intermediate_size)
    IV = ref_shim.apply(clipcfg)
    bs = sys.modules["model.segment_anything.build_sam"]

    def small_sam(checkpoint=None):
        return bs._build_sam(encoder_embed_dim=cfg.sam_embed_dim, encoder_depth=cfg.sam_depth,
                             encoder_num_heads=cfg.sam_num_heads,
                             encoder_global_attn_indexes=list(cfg.sam_global_attn_indexes), checkpoint=None)

    IV.build_sam_vit_h = small_sam
    from model.llava.model.language_model.llava_llama import LlavaConfig

    hc = LlavaConfig(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                     num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                     num_key_value_heads=cfg.num_attention_heads, vocab_size=cfg.vocab_size,
                     rms_norm_eps=cfg.rms_norm_eps, max_position_embeddings=cfg.max_position_embeddings,
                     bos_token_id=cfg.bos_token_id, eos_token_id=cfg.eos_token_id, pad_token_id=cfg.pad_token_id)
    hc._attn_implementation = "eager"
    for k, v in dict(use_fusion=False, use_uncertainty=False, token_type=cfg.token_type, train_mask_decoder=True,
                     out_dim=cfg.out_dim, vision_tower=cfg.vision_tower, mm_vision_tower=cfg.vision_tower,
                     mm_vision_select_layer=cfg.mm_vision_select_layer, mm_vision_select_feature="patch",
                     pretrain_mm_mlp_adapter=None, mm_hidden_size=cfg.clip_hidden_size, mm_use_im_start_end=True,
                     img_emb_len=cfg.img_emb_len, seg_token_idx=cfg.seg_token_idx, hseg_token_idx=None,
                     oseg_token_idx=None, hC_sam_view_type=cfg.hC_sam_view_type, oC_sam_view_type=cfg.oC_sam_view_type,
                     hC_loss_weight=cfg.hC_loss_weight, oC_loss_weight=cfg.oC_loss_weight,
                     multiview_channels=cfg.multiview_channels, multiview_cam_cond=cfg.multiview_cam_cond,
                     cam_encoder_type=cfg.cam_encoder_type, hC_question_type="simple", oC_question_type="simple").items():
        setattr(hc, k, v)
    m = IV.InteractVLMForCausalLM(hc)
    m.get_model().initialize_vision_modules(hc)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    allowed = ("point_embeddings", "not_a_point_embed", "mask_downscaling", "iou_prediction_head", "rotary_emb")
    assert all(any(a in k for a in allowed) for k in missing), [k for k in missing if not any(a in k for a in allowed)]
    return m.eval()


def install_generate_shim(m, scripted):
    """HF 5.x generate() no longer fits the reference's 4.31 contract; replace it with the loop HF 4.31 greedy search
    performs when use_cache=False: one LlavaLlamaForCausalLM.forward over the whole sequence per step."""
    from model.llava.model.language_model.llava_llama import LlavaLlamaForCausalLM

    def generate(self, images=None, input_ids=None, max_new_tokens=32, **kw):
        ids = input_ids
        done = torch.zeros(ids.shape[0], dtype=torch.bool)
        last = None
        for step in range(max_new_tokens):
            out = LlavaLlamaForCausalLM.forward(self, input_ids=ids, attention_mask=torch.ones_like(ids), images=images,
                                                past_key_values=None, use_cache=False, output_hidden_states=True,
                                                return_dict=True)
            last = out.hidden_states
            nxt = out.logits[:, -1].float().argmax(-1)
            if scripted is not None:
                nxt = scripted[:, step]
            nxt = torch.where(done, torch.full_like(nxt, self.config.pad_token_id), nxt)
            ids = torch.cat([ids, nxt[:, None]], 1)
            done |= nxt == self.config.eos_token_id
            if bool(done.all()):
                break
        return types.SimpleNamespace(sequences=ids, hidden_states=(last,))

    m.generate = types.MethodType(generate, m)


def run_reference(m, cfg, ids, ans, clip, sam, cam, dtype):
    B = ids.shape[0]
    full_ids = torch.cat([ids, ans], 1)
    size = (cfg.sam_img_size, cfg.sam_img_size)
    stages = {}
    hooks = [
        m.get_model().mm_projector.register_forward_hook(lambda mod, i, o: stages.__setitem__("clip_proj", o.detach().float())),
        m.get_model().text_hidden_fcs[0].register_forward_hook(lambda mod, i, o: stages.__setitem__("fcs_all", o.detach().float())),
        m.get_model().visual_model.image_encoder.register_forward_hook(
            lambda mod, i, o: stages.setdefault("sam_emb", []).append(o.detach().float())),
        m.get_model().visual_model.mask_decoder.register_forward_hook(
            lambda mod, i, o: stages.setdefault("low_res", []).append(o[0].detach().float())),
        m.get_model().norm.register_forward_hook(lambda mod, i, o: stages.__setitem__("hidden", o.detach().float())),
    ]
    with torch.no_grad():
        # teacher-forced path; images_clip must have batch 1 on this path (InteractVLM.py:346)
        fw = []
        for b in range(B):
            fw.append(m(images=sam[b:b + 1].to(dtype), images_clip=clip[b:b + 1].to(dtype), input_ids=full_ids[b:b + 1],
                        labels=full_ids[b:b + 1], attention_masks=torch.ones_like(full_ids[b:b + 1]),
                        offset=torch.tensor([0, 1]), masks_list=[torch.zeros(cfg.multiview_channels, 1, *size)],
                        label_list=[torch.zeros(size)], gt_contact_3d_list=[None], cam_params=cam[b:b + 1].to(dtype),
                        resize_list=[size], ds_name_list=["damon_hcontact"], mask_paths_list=[None], inference=True))
        st_forward = {k: (v if not isinstance(v, list) else list(v)) for k, v in stages.items()}
        stages.clear()
        install_generate_shim(m, ans)
        ev = m.evaluate(clip.to(dtype), sam.to(dtype), ids, cam.to(dtype), [size] * B, [size] * B,
                        contact_type="hcontact", max_new_tokens=ans.shape[1])
        st_eval = dict(stages)
        stages.clear()
        install_generate_shim(m, None)
        free = m.generate(images=clip.to(dtype), input_ids=ids, max_new_tokens=4)
    for h in hooks:
        h.remove()
    return fw, st_forward, ev, st_eval, free.sequences[:, ids.shape[1]:]


def gold_tiny():
    cfg = IVLMConfig.tiny()
    sd = S.make_state_dict(cfg, seed=TINY_SEED["weights"])
    ids, ans, clip, sam, cam = tiny_inputs(cfg, batch=1)
    m = build_reference(cfg, sd)
    out = {}
    for tag, dtype in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        mm = m.float() if dtype == torch.float32 else m.bfloat16()
        fw, stf, ev, ste, free = run_reference(mm, cfg, ids, ans, clip, sam, cam, dtype)
        pm_f = torch.stack([r["pred_masks"][0] for r in fw], 0).float()
        pm_e = torch.stack(ev["pred_masks"], 0).float()
        c_f = torch.cat([r["pred_human_3d_contact"] for r in fw], 0).float()
        c_e = ev["pred_contact_3d"].float()
        # last-step hidden states of a no-cache generate == one causal forward over output_ids[:, :-1]
        print(tag, "evaluate vs model_forward: masks", (pm_f - pm_e).abs().max().item(), "contact",
              (c_f - c_e).abs().max().item())
        assert torch.equal(ev["output_ids"], torch.cat([ids, ans], 1))
        seg_row = cfg.img_emb_len + ids.shape[1] + (ans.shape[1] - 3) - 1
        out[f"{tag}_clip_proj"] = ste["clip_proj"][:, ::EMB_STRIDE, ::EMB_STRIDE].numpy()
        out[f"{tag}_hidden_seg"] = ste["hidden"][:, seg_row].numpy()
        out[f"{tag}_pred_emb"] = ste["fcs_all"][:, seg_row].numpy()
        out[f"{tag}_sam_emb"] = torch.stack(ste["sam_emb"], 0)[:, :, ::EMB_STRIDE, ::EMB_STRIDE, ::EMB_STRIDE].numpy()
        out[f"{tag}_low_res"] = torch.stack(ste["low_res"], 0)[:, :, 0, ::LOW_STRIDE, ::LOW_STRIDE].numpy()
        out[f"{tag}_pred_masks"] = pm_e[:, :, ::FULL_STRIDE, ::FULL_STRIDE].numpy()
        out[f"{tag}_contact"] = c_e.numpy()
        out[f"{tag}_contact_forward"] = c_f.numpy()
        out[f"{tag}_greedy4"] = free.numpy()
        print(tag, "logit std", pm_e.std().item(), "contact mean", c_e.mean().item(), "frac>=0.5",
              (c_e >= 0.5).float().mean().item(), "greedy", free.tolist())
    np.savez_compressed(GOLD / "tiny_model.npz", **out)
    print("tiny goldens:", {k: v.shape for k, v in out.items()})


def gold_decoder():
    """Full-size mask-decoder tail (the decoder has no size knobs): reference prompt encoder + mask decoder +
    postprocess on seeded embeddings."""
    cfg = IVLMConfig.tiny()
    sd = S.make_state_dict(cfg, seed=TINY_SEED["weights"])
    m = build_reference(cfg, sd).float()
    rng = np.random.default_rng(77)
    emb = torch.from_numpy(rng.standard_normal((4, 256, 64, 64), dtype=np.float32)).bfloat16().float()
    prompt = torch.from_numpy(rng.standard_normal((1, 4, 256), dtype=np.float32) * 0.5).bfloat16().float()
    vm = m.get_model().visual_model
    with torch.no_grad():
        sparse, dense = vm.prompt_encoder(points=None, boxes=None, masks=None, text_embeds=prompt)
        low, _ = vm(image_embeddings=emb, llava_features=None, sparse_prompt_embeddings=sparse,
                    dense_prompt_embeddings=dense, ds_name="hcontact")
        full = vm.postprocess_masks(low, input_size=(1024, 1024), original_size=(1024, 1024))
        pe = vm.prompt_encoder.get_dense_pe()
    np.savez_compressed(GOLD / "decoder.npz", low_res=low[:, 0, ::2, ::2].numpy(), full=full[:, 0, ::FULL_STRIDE, ::FULL_STRIDE].numpy(),
                        dense_pe=pe[0, :, ::4, ::4].numpy())
    print("decoder goldens: low-res std", low.std().item())
