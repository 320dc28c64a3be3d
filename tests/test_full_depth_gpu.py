"""Full-depth parity at the BASELINE configuration (LLaMA-2-13B: 40 layers, SAM ViT-H: 32 blocks, CLIP-L: 23 layers, seeded
random-init weights) on a B200: model.evaluate() of the CUDA path at batch 1 and batch 8 against the oracle restatement
(oracle/model.py) executed with stock torch ops on the GPU -- in fp32 (the "exact" answer) and in bf16 (what the eager bf16
reference computes, evaluate.py:532; its distance from the fp32 run is the yardstick, as in the tiny-config test).

What the north star's tolerances mean here (also stated in DESIGN.md section 2): the 1e-3 max-abs target on the contact
probabilities holds for the fp32 tail (upsample + lift) on shared inputs (test_model_gpu.py::test_tail_given_identical_inputs_is_1e3)
and CANNOT hold end to end in bf16 -- the reference's own bf16 run differs from its fp32 run by more than that -- so end to end
the bar is "no further from the fp32 answer than 3x the reference's own bf16 deviation", an identical contact vertex set away
from the threshold, and an agreement F1 within 0.5 pt of the reference-bf16 F1."""
import math

import numpy as np
import pytest
import torch

from interactvlm_b200 import synthetic as S
from interactvlm_b200.config import IVLMConfig
from oracle import lift as OL
from oracle import model as OM

pytestmark = pytest.mark.gpu
SIZE = (1024, 1024)
N_PRE, N_POST, N_ANS = 40, 30, 24   # bench.py's prompt


def _inputs(cfg, batch, seed):
    ids, ans = S.make_prompt_ids(cfg, batch, n_pre=N_PRE, n_post=N_POST, n_answer=N_ANS, seed=seed)
    clip, sam = S.make_images(cfg, batch, seed=seed)
    cam = torch.from_numpy(np.broadcast_to(S.HCONTACT_CAM_PARAMS, (batch, cfg.multiview_channels, 5)).copy()).bfloat16()
    return torch.from_numpy(ids), torch.from_numpy(ans), torch.from_numpy(clip).bfloat16(), torch.from_numpy(sam).bfloat16(), cam


def _oracle(sd, cfg, dtype, ids_full, clip, sam, cam, maps, trace=None):
    """model_forward(inference=True) of the oracle on the GPU = evaluate() downstream of token selection (the
    reference's two entry points agree bit for bit on scripted tokens).  Returns (pred_masks [B][V,H,W] fp32 cpu, contact, low-res)."""
    w = OM.W(sd, dtype, device="cuda")
    B = ids_full.shape[0]
    with torch.no_grad():
        feats = OM.encode_images(w, cfg, clip.to(dtype))
        emb = OM.splice_embeddings(w, cfg, ids_full, feats)
        hidden = OM.llama_forward(w, cfg, emb, trace=None if trace is None else trace.setdefault("llm", []))
        st = {}
        masks = []
        rows, tokens = OM.seg_rows(cfg, ids_full, with_tokens=True)
        lows = []
        for b in range(B):
            e = OM.sam_image_encoder(w, cfg, sam[b].to(dtype), trace=None if (trace is None or b > 0) else trace.setdefault("sam", []))
            pe = OM.text_hidden_fcs(w, hidden[b, rows[b]])
            prompt = OM.process_embeddings(w, cfg, pe, cam[b].to(dtype), tokens[b])
            low = OM.mask_decoder(w, cfg, e, prompt)
            lows.append(low.float().cpu())
            masks.append(OM.postprocess_masks(cfg, low, SIZE, SIZE)[:, 0].float().cpu())
    p2v, bary, n = maps
    contact = OL.lift_human(np.stack([m.numpy() for m in masks], 0), p2v, bary, n)
    del w
    torch.cuda.empty_cache()
    return masks, contact, torch.stack(lows, 0)


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


@pytest.fixture(scope="module")
def full(ctx):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = IVLMConfig.full()
    sd = S.make_state_dict(cfg, seed=11, device="cuda", gain=0.5)
    maps = S.make_mesh_lift_maps(seed=3) + (S.N_SMPL,)
    # give the random-init mask logits a spread of a few units (otherwise every probability is ~0.5 and the
    # vertex-set checks are vacuous): scale the last hypernetwork layer by a power of two chosen from one fp32 oracle pass
    ids, ans, clip, sam, cam = _inputs(cfg, 1, seed=100)
    _, _, low = _oracle(sd, cfg, torch.float32, torch.cat([ids, ans], 1), clip.cuda(), sam.cuda(), cam.cuda(), maps)
    std = float(low.std())
    f = 2.0 ** round(math.log2(4.0 / max(std, 1e-6)))
    hp = S.SAM_PREFIX + "mask_decoder.output_hypernetworks_mlps.0.layers.2."
    sd[hp + "weight"].mul_(f)
    sd[hp + "bias"].mul_(f)
    print(f"low-res logit std {std:.4f} -> hypernetwork output scaled by {f}")
    from interactvlm_b200.model import InteractVLMForCausalLM

    model = InteractVLMForCausalLM(cfg, sd, ctx=ctx)
    model.set_human_lift_maps(maps[0], maps[1])
    return cfg, sd, model, maps


@pytest.mark.parametrize("batch", [1, 8])
def test_full_depth_evaluate_vs_oracle(full, batch):
    cfg, sd, model, maps = full
    ids, ans, clip, sam, cam = _inputs(cfg, batch, seed=100)
    ids_full = torch.cat([ids, ans], 1)
    tr32, trp = ({}, {}) if batch == 1 else (None, None)
    m32, c32, low32 = _oracle(sd, cfg, torch.float32, ids_full, clip.cuda(), sam.cuda(), cam.cuda(), maps, trace=tr32)
    m16, c16, _ = _oracle(sd, cfg, torch.bfloat16, ids_full, clip.cuda(), sam.cuda(), cam.cuda(), maps)
    model.eng.trace = trp
    try:
        out = model.evaluate(clip, sam, ids, cam, [SIZE] * batch, [SIZE] * batch, contact_type="hcontact", max_new_tokens=N_ANS,
                             scripted=ans)
    finally:
        model.eng.trace = None
    assert torch.equal(out["output_ids"].cpu(), ids_full)                      # token bookkeeping: exact
    c = out["pred_contact_3d"].cpu().numpy()
    scale = max(float(low32.abs().max()), 1e-6)
    ref_mask_noise = max(float((a - b).abs().max()) for a, b in zip(m16, m32))
    mask_err = max(float((out["pred_masks"][b].cpu() - m32[b]).abs().max()) for b in range(batch))
    ref_noise = float(np.abs(c16 - c32).max())
    err = float(np.abs(c - c32).max())
    print(f"[B={batch}] mask logits (scale {scale:.2f}, std {float(low32.std()):.2f}): ours vs fp32 {mask_err:.4f}; oracle bf16 vs fp32 {ref_mask_noise:.4f}")
    print(f"[B={batch}] contact: ours vs fp32 {err:.5f}; oracle bf16 vs fp32 {ref_noise:.5f}; 1e-3 end-to-end target met: {err <= 1e-3}")
    assert mask_err <= max(3 * ref_mask_noise, 0.03 * scale)
    assert err <= max(3 * ref_noise, 0.02)
    far = np.abs(c32 - 0.5) > max(3 * ref_noise, 0.02)
    assert far.mean() > 0.3                                                     # the check is not vacuous
    assert np.array_equal((c >= 0.5)[far], (c32 >= 0.5)[far])                   # contact vertex set away from the threshold
    gt = (c32 >= 0.5).astype(np.float32)
    f1, _, _ = OL.f1_metrics(c, gt)
    f1_ref, _, _ = OL.f1_metrics(c16, gt)
    print(f"[B={batch}] agreement F1 with the fp32 contact set: ours {f1:.4f}, oracle bf16 {f1_ref:.4f}; contact fraction {gt.mean():.3f}")
    # "within 0.5 pt of the reference": both runs sit at bf16 noise from the fp32 set, where single vertices next to the threshold
    # flip; one vertex is worth ~1/n_pos of F1, so with few contact vertices (batch 1: ~280) the bar cannot be finer than a few
    # vertices -- 0.5 pt, or three vertices' worth where that is more
    assert f1 >= f1_ref - max(0.005, 3.0 / max(float(gt.sum()), 1.0))
    if batch == 1:   # error growth over depth (printed for the record, bounded loosely)
        sam_curve = [_rel(a.reshape(-1, a.shape[-1]), b.reshape(-1, b.shape[-1])) for a, b in zip(trp["sam"], tr32["sam"])]
        S_p = ids.shape[1] - 1 + cfg.clip_tokens - 1   # our trace covers the prefill rows; causal: same rows of the oracle's full pass
        llm_curve = [_rel(a.reshape(-1, a.shape[-1]), b[:, :S_p].reshape(-1, b.shape[-1]))
                     for a, b in zip(trp["llm"][:cfg.num_hidden_layers], tr32["llm"])]
        print("SAM residual stream rel. error after blocks 1..32:", " ".join(f"{x:.4f}" for x in sam_curve))
        print("LLaMA residual stream rel. error after layers 1..40:", " ".join(f"{x:.4f}" for x in llm_curve))
        assert len(sam_curve) == cfg.sam_depth and len(llm_curve) == cfg.num_hidden_layers
        assert max(sam_curve) < 0.05 and max(llm_curve) < 0.05


def test_full_depth_greedy_tokens(full):
    """Greedy decoding through the paged KV cache at full depth: the tokens must be the argmax of the fp32 oracle's logits
    wherever the oracle's top-2 margin exceeds the bf16 noise (random-init logits are nearly flat, so margins are reported)."""
    cfg, sd, model, maps = full
    ids, ans, clip, sam, cam = _inputs(cfg, 2, seed=300)
    G = 4
    out_ids, _ = model.generate(clip, ids, max_new_tokens=G)
    w = OM.W(sd, torch.float32, device="cuda")
    with torch.no_grad():
        seq = out_ids[:, :ids.shape[1] + G - 1]                     # teacher-force OUR tokens through the oracle
        hid = OM.lm_hidden(w, cfg, clip.cuda().float(), seq)
        L = ids.shape[1]
        rows = [L - 1 + cfg.img_emb_len + t for t in range(G)]
        logits = OM.lm_logits(w, hid[:, rows]).float()
    top2 = logits.topk(2, -1)
    margin = (top2.values[..., 0] - top2.values[..., 1]).cpu()
    want = top2.indices[..., 0].cpu()
    got = out_ids[:, L:L + G]
    print("oracle top-2 margins:", margin.tolist(), "agree:", (got == want).tolist())
    clear = margin > 0.05 * logits.abs().max().item()
    assert bool(((got == want) | ~clear).all())
