"""CPU (kernel emulator): the caller-side pieces around the hot path -- GPU input preparation semantics, convert_contacts,
on-disk result formats of run_demo.py, metrics and the batched validate loop."""
import numpy as np
import pytest
import torch

from interactvlm_b200 import harness as Hn
from interactvlm_b200 import synthetic as S
from interactvlm_b200.config import IVLMConfig
from interactvlm_b200.model import InteractVLMForCausalLM
from oracle import lift as OL
from oracle import model as OM
from oracle.make_goldens_model import TINY_SEED, tiny_inputs

from emu import EmuContext

SIZE = (1024, 1024)


@pytest.fixture(scope="module")
def model():
    cfg = IVLMConfig.tiny()
    sd = S.make_state_dict(cfg, seed=TINY_SEED["weights"])
    m = InteractVLMForCausalLM(cfg, sd, ctx=EmuContext())
    p2v, bary = S.make_mesh_lift_maps(seed=TINY_SEED["maps"])
    m.set_human_lift_maps(p2v, bary)
    return m


def test_prepare_inputs_matches_reference_preprocess(model):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (1, 224, 224, 3), dtype=np.uint8)
    views = rng.integers(0, 256, (1, 4, 1024, 768, 3), dtype=np.uint8)   # a non-square view: zero padded on the right
    clip, sam, resize = Hn.prepare_inputs(model, img, views)
    assert clip.shape == (1, 3, 224, 224) and sam.shape == (1, 4, 3, 1024, 1024) and resize == [(1024, 768)]
    # run_demo.py:65-79 preprocess(): (x - mean) / std, pad right/bottom with zeros
    x = torch.from_numpy(views[0, 2]).permute(2, 0, 1).float()
    ref = (x - torch.tensor([123.675, 116.28, 103.53]).view(-1, 1, 1)) / torch.tensor([58.395, 57.12, 57.375]).view(-1, 1, 1)
    ref = torch.nn.functional.pad(ref, (0, 1024 - 768, 0, 0)).bfloat16()
    assert torch.equal(sam[0, 2].cpu(), ref)
    c = torch.from_numpy(img[0]).permute(2, 0, 1).float() / 255.0
    cref = ((c - torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(-1, 1, 1))
            / torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(-1, 1, 1))
    assert (clip[0].float().cpu() - cref).abs().max().item() < 2e-2  # bf16 storage of values up to ~2.2


def test_convert_contacts_and_npz_format(tmp_path, model):
    mapping = S.make_smplx_matrix(seed=0)
    conv = Hn.ContactConverter(model, mapping)
    contact = torch.rand(1, S.N_SMPL)
    smplx = conv(contact)
    assert smplx.shape == (S.N_SMPLX,)   # the reference's squeeze() drops the batch of 1
    assert np.abs(smplx.numpy() - OL.convert_contacts(contact.numpy(), mapping)[0]).max() < 1e-5
    f = Hn.save_hcontact(tmp_path / "img0", contact, smplx)
    z = np.load(f)
    assert f.name == "img0_hcontact_vertices.npz" and set(z.files) == {"pred_contact_3d_smplh", "pred_contact_3d_smplx"}
    assert z["pred_contact_3d_smplh"].shape == (1, S.N_SMPL) and z["pred_contact_3d_smplx"].shape == (S.N_SMPLX,)
    f = Hn.save_ocontact(tmp_path / "img0", torch.rand(1, 1234))
    assert f.name == "img0_oafford_vertices.npz" and np.load(f)["pred_contact_3d"].shape == (1, 1234)


def test_metrics_match_reference_formula():
    gt = (torch.rand(3, 100) > 0.6).float()
    pred = torch.rand(3, 100)
    f1, p, r = Hn.h_contact_metrics(gt, pred)
    f1o, po, ro = OL.f1_metrics(pred.numpy(), gt.numpy())
    assert abs(f1 - f1o) < 1e-6 and abs(p - po) < 1e-6 and abs(r - ro) < 1e-6


def test_validate_batches_and_scores(model):
    cfg = model.config
    ids, ans, clip, sam, cam = tiny_inputs(cfg, 2)
    ref = model.evaluate(clip, sam, ids, cam, [SIZE] * 2, [SIZE] * 2, max_new_tokens=ans.shape[1], scripted=ans)
    samples = [dict(images_clip=clip[b], images=sam[b], input_ids=ids[b], cam_params=cam[b], resize=SIZE, original_size=SIZE,
                    scripted=ans[b], gt_contact_3d=(ref["pred_contact_3d"][b] >= 0.5).float()) for b in range(2)]
    preds, metrics = Hn.validate(model, samples, batch_size=2, max_new_tokens=ans.shape[1])
    assert preds.shape == (2, S.N_SMPL) and torch.allclose(preds, ref["pred_contact_3d"], atol=1e-6)
    assert metrics["n"] == 2 and metrics["f1"] > 0.999
    preds1, _ = Hn.validate(model, samples, batch_size=1, max_new_tokens=ans.shape[1])
    assert (preds1 - preds).abs().max().item() < 2e-2


def test_prompt_and_camera_helpers_match_reference_goldens():
    """build_prompt / tokenizer_image_token / normalize_cam_params against outputs of the reference's own functions
    (tests/golden/prompt.json, written by oracle/make_goldens_prompt.py)."""
    import json
    from pathlib import Path

    import torch

    from interactvlm_b200 import harness as Hn
    from oracle.make_goldens_prompt import CAMS, NoBosTokenizer, WordTokenizer

    gold = json.loads((Path(__file__).parent / "golden" / "prompt.json").read_text())
    assert len(gold["prompts"]) == 8
    for g in gold["prompts"]:
        text = Hn.build_prompt(g["question"], g["conv_type"], g["use_mm_start_end"])
        assert text == g["prompt"]
        assert Hn.tokenizer_image_token(text, WordTokenizer()) == g["ids_bos"]
        ids = Hn.tokenizer_image_token(text, NoBosTokenizer(), return_tensors="pt")
        assert ids.dtype == torch.long and ids.tolist() == g["ids_nobos"]
    for cam, want in zip(CAMS, gold["cams"]):
        assert Hn.normalize_cam_params(cam).tolist() == want
    from interactvlm_b200 import synthetic as S

    assert torch.allclose(Hn.human_cam_params("4MV-Z_Vitru")[0], torch.from_numpy(S.HCONTACT_CAM_PARAMS).float(), atol=1e-7)
