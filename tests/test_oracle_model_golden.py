"""CPU: the torch-CPU oracle of the model stages (oracle/model.py) against golden vectors recorded from the
UNMODIFIED reference (oracle/make_goldens_model.py: reference InteractVLMForCausalLM.evaluate / model_forward,
prompt encoder + mask decoder + postprocess) on the same seeded weights and inputs."""
from pathlib import Path

import numpy as np
import pytest
import torch

from interactvlm_b200 import synthetic as S
from interactvlm_b200.config import IVLMConfig
from oracle import model as OM
from oracle.make_goldens_model import EMB_STRIDE, FULL_STRIDE, LOW_STRIDE, TINY_SEED, tiny_inputs

GOLD = Path(__file__).parent / "golden"
SIZE = (1024, 1024)


@pytest.fixture(scope="module")
def tiny():
    cfg = IVLMConfig.tiny()
    sd = S.make_state_dict(cfg, seed=TINY_SEED["weights"])
    return cfg, sd, tiny_inputs(cfg, 1), np.load(GOLD / "tiny_model.npz")


def test_oracle_evaluate_matches_reference_fp32(tiny):
    cfg, sd, (ids, ans, clip, sam, cam), gold = tiny
    p2v, bary = S.make_mesh_lift_maps(seed=TINY_SEED["maps"])
    st = {}
    out = OM.evaluate(sd, cfg, clip, sam, ids, cam, [SIZE], [SIZE], lift_maps=(p2v, bary, S.N_SMPL),
                      max_new_tokens=ans.shape[1], scripted=ans, dtype=torch.float32, stages=st)
    assert torch.equal(out["output_ids"], torch.cat([ids, ans], 1))
    seg_row = cfg.img_emb_len + ids.shape[1] + (ans.shape[1] - 3) - 1
    assert OM.seg_rows(cfg, out["output_ids"]) == [[seg_row]]

    def close(name, a, tol):
        g = gold["f32_" + name]
        err = np.abs(a.float().numpy() - g).max() / max(np.abs(g).max(), 1e-6)
        assert err < tol, (name, err)

    # fp32 vs fp32: only summation-order noise (different BLAS blocking / conv lowering)
    close("hidden_seg", st["hidden"][:, seg_row], 1e-5)
    close("pred_emb", st["pred_embeddings"][0], 1e-5)
    close("sam_emb", torch.stack(st["sam_embeddings"], 0)[:, :, ::EMB_STRIDE, ::EMB_STRIDE, ::EMB_STRIDE], 1e-5)
    close("low_res", torch.stack(st["low_res"], 0)[:, :, 0, ::LOW_STRIDE, ::LOW_STRIDE], 2e-5)
    close("pred_masks", torch.stack(out["pred_masks"], 0)[:, :, ::FULL_STRIDE, ::FULL_STRIDE], 2e-5)
    c, g = out["pred_contact_3d"].numpy(), gold["f32_contact"]
    assert np.abs(c - g).max() < 1e-5
    near = np.abs(g - 0.5) < 1e-4
    assert np.array_equal((c >= 0.5)[~near], (g >= 0.5)[~near])  # contact vertex set, ties excluded
    assert near.sum() <= 2 and 0.05 < (g >= 0.5).mean() < 0.95
    # the reference's evaluate() and model_forward(inference=True) agree bit-for-bit
    assert np.array_equal(gold["f32_contact"], gold["f32_contact_forward"])


def test_oracle_greedy_tokens_match_reference(tiny):
    cfg, sd, (ids, ans, clip, sam, cam), gold = tiny
    w = OM.W(sd)
    with torch.no_grad():
        seq, _, greedy = OM.greedy_generate(w, cfg, clip, ids, 4)
    assert greedy.tolist() == gold["f32_greedy4"].tolist()
    assert seq[:, ids.shape[1]:].tolist() == gold["f32_greedy4"].tolist()


def test_oracle_model_forward_equals_evaluate(tiny):
    cfg, sd, (ids, ans, clip, sam, cam), gold = tiny
    full = torch.cat([ids, ans], 1)
    out = OM.model_forward(sd, cfg, sam, clip, full, cam, [SIZE], [SIZE])
    g = gold["f32_pred_masks"]
    a = torch.stack(out["pred_masks"], 0)[:, :, ::FULL_STRIDE, ::FULL_STRIDE].numpy()
    assert np.abs(a - g).max() / np.abs(g).max() < 2e-5


def test_oracle_bf16_mode_is_at_reference_bf16_noise(tiny):
    """Running the oracle in bfloat16 lands as close to the fp32 reference as the reference's own bf16 run does
    (same order of magnitude); this is the noise floor the GPU tests' tolerances are derived from."""
    cfg, sd, (ids, ans, clip, sam, cam), gold = tiny
    out = OM.evaluate(sd, cfg, clip, sam, ids, cam, [SIZE], [SIZE], max_new_tokens=ans.shape[1], scripted=ans,
                      dtype=torch.bfloat16)
    a = torch.stack(out["pred_masks"], 0)[:, :, ::FULL_STRIDE, ::FULL_STRIDE].numpy()
    ref_noise = np.abs(gold["bf16_pred_masks"] - gold["f32_pred_masks"]).max()
    ours = np.abs(a - gold["f32_pred_masks"]).max()
    assert ours < 4 * ref_noise + 1e-3, (ours, ref_noise)


def test_oracle_decoder_matches_reference():
    cfg = IVLMConfig.tiny()
    sd = S.make_state_dict(cfg, seed=TINY_SEED["weights"])
    gold = np.load(GOLD / "decoder.npz")
    rng = np.random.default_rng(77)
    emb = torch.from_numpy(rng.standard_normal((4, 256, 64, 64), dtype=np.float32)).bfloat16().float()
    prompt = torch.from_numpy(rng.standard_normal((1, 4, 256), dtype=np.float32) * 0.5).bfloat16().float()
    w = OM.W(sd)
    with torch.no_grad():
        low = OM.mask_decoder(w, cfg, emb, prompt)
        full = OM.postprocess_masks(cfg, low, SIZE, SIZE)
        pe = OM.dense_pe(w, cfg)
    assert np.abs(pe[0, :, ::4, ::4].numpy() - gold["dense_pe"]).max() < 1e-5
    assert np.abs(low[:, 0, ::2, ::2].numpy() - gold["low_res"]).max() / np.abs(gold["low_res"]).max() < 2e-5
    assert np.abs(full[:, 0, ::FULL_STRIDE, ::FULL_STRIDE].numpy() - gold["full"]).max() / np.abs(gold["full"]).max() < 2e-5
