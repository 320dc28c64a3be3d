"""CPU: host logic of interactvlm_b200/model.py (stage order, weight re-layout, window row maps, KV-cache and [SEG]
bookkeeping, API surface) with the kernels replaced by tests/emu.py, against the oracle pinned to the reference.
The numbers here are NOT the product's (the CUDA kernels are checked by the -m gpu tests); the structure is."""
import numpy as np
import pytest
import torch

from interactvlm_b200 import synthetic as S
from interactvlm_b200.config import IVLMConfig
from interactvlm_b200.model import InteractVLMForCausalLM, config_from_hf, save_pretrained
from oracle import model as OM
from oracle.make_goldens_model import TINY_SEED, tiny_inputs

from emu import EmuContext

SIZE = (1024, 1024)


@pytest.fixture(scope="module")
def setup():
    cfg = IVLMConfig.tiny()
    sd = S.make_state_dict(cfg, seed=TINY_SEED["weights"])
    model = InteractVLMForCausalLM(cfg, sd, ctx=EmuContext())
    p2v, bary = S.make_mesh_lift_maps(seed=TINY_SEED["maps"])
    model.set_human_lift_maps(p2v, bary)
    return cfg, sd, model, (p2v, bary)


def test_evaluate_structure_matches_oracle(setup):
    cfg, sd, model, (p2v, bary) = setup
    ids, ans, clip, sam, cam = tiny_inputs(cfg, 1)
    out = model.evaluate(clip, sam, ids, cam, [SIZE], [SIZE], contact_type="hcontact", max_new_tokens=ans.shape[1],
                         scripted=ans)
    ref = OM.evaluate(sd, cfg, clip, sam, ids, cam, [SIZE], [SIZE], lift_maps=(p2v, bary, S.N_SMPL),
                      max_new_tokens=ans.shape[1], scripted=ans, dtype=torch.float32)
    assert torch.equal(out["output_ids"].cpu(), ref["output_ids"])
    pm, rm = out["pred_masks"][0], ref["pred_masks"][0]
    assert pm.shape == rm.shape == (4, 1024, 1024) and pm.dtype == torch.float32
    # bf16 storage between stages: a few percent of the logit scale, far below any structural error (O(scale))
    assert (pm - rm).abs().max().item() < 0.08 * rm.abs().max().item()
    c, rc = out["pred_contact_3d"].numpy(), ref["pred_contact_3d"].numpy()
    assert c.shape == (1, S.N_SMPL)
    assert np.abs(c - rc).max() < 0.08
    assert ((c >= 0.5) == (rc >= 0.5)).mean() > 0.97


def test_greedy_decode_through_kv_cache_matches_oracle(setup):
    cfg, sd, model, _ = setup
    ids, ans, clip, sam, cam = tiny_inputs(cfg, 1)
    w = OM.W(sd)
    with torch.no_grad():
        seq, hidden, _ = OM.greedy_generate(w, cfg, clip, ids, 4)
    out_ids, hid = model.generate(clip, ids, max_new_tokens=4)
    assert out_ids.tolist() == seq.tolist()
    n = hidden.shape[1]
    err = (hid[:, :n].float() - hidden).abs().max().item()
    assert err < 0.05 * hidden.abs().max().item(), err


def test_batch_of_two_and_missing_seg(setup):
    cfg, sd, model, (p2v, bary) = setup
    ids, ans, clip, sam, cam = tiny_inputs(cfg, 2)
    ans = ans.clone()
    ans[1, -3] = 5  # second sample never emits [SEG]
    # batched extension of the reference's batch-1 rule (InteractVLM.py:596-601,618): a sample without [SEG] keeps an empty
    # [0,H,W] mask stack and gets a zero contact row; the other samples are lifted as usual
    out0 = model.evaluate(clip, sam, ids, cam, [SIZE] * 2, [SIZE] * 2, max_new_tokens=ans.shape[1], scripted=ans)
    assert out0["pred_masks"][1].shape[0] == 0 and out0["pred_masks"][0].shape[0] == cfg.multiview_channels
    assert out0["pred_contact_3d"].shape == (2, S.N_SMPL) and float(out0["pred_contact_3d"][1].abs().max()) == 0.0
    # no sample with a [SEG] at all: None, like the reference
    ans0 = ans.clone()
    ans0[0, -3] = 5
    assert model.evaluate(clip, sam, ids, cam, [SIZE] * 2, [SIZE] * 2, max_new_tokens=ans.shape[1], scripted=ans0)["pred_contact_3d"] is None
    ans[1, -3] = cfg.seg_token_idx
    out = model.evaluate(clip, sam, ids, cam, [SIZE] * 2, [SIZE] * 2, max_new_tokens=ans.shape[1], scripted=ans)
    ref = OM.evaluate(sd, cfg, clip, sam, ids, cam, [SIZE] * 2, [SIZE] * 2, lift_maps=(p2v, bary, S.N_SMPL),
                      max_new_tokens=ans.shape[1], scripted=ans)
    assert out["pred_contact_3d"].shape == (2, S.N_SMPL)
    assert np.abs(out["pred_contact_3d"].numpy() - ref["pred_contact_3d"].numpy()).max() < 0.08
    assert torch.equal(out["pred_contact_3d"][0], out0["pred_contact_3d"][0])   # sample 0 does not depend on its neighbour


def test_view_cache_bookkeeping(setup):
    """Host logic of the encoder view cache (which views are encoded, where their embeddings are filed, how the batch is
    re-assembled) with the emulated kernels: identical outputs, repeated views encoded once."""
    cfg, sd, model, _ = setup
    ids, ans, clip, sam, cam = tiny_inputs(cfg, 3)
    sam = sam.clone()
    sam[1] = sam[0]
    sam[2, 0] = sam[0, 2]
    args = (clip, sam, ids, cam, [SIZE] * 3, [SIZE] * 3)
    base = model.evaluate(*args, max_new_tokens=ans.shape[1], scripted=ans)
    encoded = []
    orig = model.eng.sam_encode
    model.eng.sam_encode = lambda x: (encoded.append(x.shape[0]), orig(x))[1]
    try:
        model.enable_view_cache(max_entries=6)
        for call in range(2):
            out = model.evaluate(*args, max_new_tokens=ans.shape[1], scripted=ans)
            assert torch.equal(out["pred_contact_3d"], base["pred_contact_3d"])
            assert all(torch.equal(a, b) for a, b in zip(out["pred_masks"], base["pred_masks"]))
        assert sum(encoded) == 7 + 1                       # 7 distinct views, 6 cached, the 7th re-encoded on the second call
        assert model._view_cache["hits"] == 11 and model._view_cache["inputs"].shape[0] == 6
        model.clear_view_cache()
        assert model._view_cache["inputs"] is None and model._view_cache["hits"] == 0
    finally:
        model.eng.sam_encode = orig
        model._view_cache = None


def test_model_forward_teacher_forced(setup):
    cfg, sd, model, (p2v, bary) = setup
    ids, ans, clip, sam, cam = tiny_inputs(cfg, 1)
    full = torch.cat([ids, ans], 1)
    out = model(images=sam, images_clip=clip, input_ids=full, labels=full, attention_masks=torch.ones_like(full),
                offset=torch.tensor([0, 1]), masks_list=[torch.zeros(4, 1, *SIZE)], label_list=[torch.zeros(SIZE)],
                gt_contact_3d_list=[None], cam_params=cam, resize_list=[SIZE], ds_name_list=["damon_hcontact"],
                mask_paths_list=[None], inference=True)
    ev = model.evaluate(clip, sam, ids, cam, [SIZE], [SIZE], max_new_tokens=ans.shape[1], scripted=ans)
    # the teacher-forced pass and the generate path see the same hidden state at the [SEG]-1 row
    assert (out["pred_masks"][0] - ev["pred_masks"][0]).abs().max().item() < 0.05 * ev["pred_masks"][0].abs().max().item()
    assert out["pred_human_3d_contact"].shape == (1, S.N_SMPL)


def test_model_forward_one_image_several_conversations(setup):
    """The reference's validation layout (InteractVLM.py:346: images_clip batch 1, offset = [0, n]): n conversation rows share one
    image -- its CLIP features, SAM embeddings, camera parameters and sizes -- and every row's [SEG] is decoded against them."""
    cfg, sd, model, _ = setup
    ids, ans, clip, sam, cam = tiny_inputs(cfg, 1)
    full = torch.cat([ids, ans], 1)
    two = torch.cat([full, full], 0)
    kw = dict(labels=two, attention_masks=torch.ones_like(two), masks_list=[torch.zeros(4, 1, *SIZE)], label_list=[torch.zeros(SIZE)],
              gt_contact_3d_list=[None], cam_params=cam, resize_list=[SIZE], ds_name_list=["damon_hcontact"], mask_paths_list=[None],
              inference=True)
    out = model(images=sam, images_clip=clip, input_ids=two, offset=torch.tensor([0, 2]), **kw)
    one = model(images=sam, images_clip=clip, input_ids=full, offset=torch.tensor([0, 1]), **{**kw, "labels": full,
                                                                                         "attention_masks": torch.ones_like(full)})
    assert len(out["pred_masks"]) == 2 and out["pred_human_3d_contact"].shape == (2, S.N_SMPL)
    for b in range(2):   # both conversations are the same prompt: same masks as the single-conversation call
        assert torch.allclose(out["pred_masks"][b], one["pred_masks"][0], atol=1e-5)
    with pytest.raises(ValueError):
        model(images=sam, images_clip=clip, input_ids=two, offset=torch.tensor([0, 1]), **kw)


def test_checkpoint_roundtrip_and_api_surface(tmp_path, setup):
    cfg, sd, model, _ = setup
    save_pretrained(tmp_path / "ckpt", cfg, sd)
    from interactvlm_b200.model import load_checkpoint_dir
    import json
    cfg2 = config_from_hf(json.loads((tmp_path / "ckpt" / "config.json").read_text()))
    assert cfg2.to_dict() == cfg.to_dict()
    sd2 = load_checkpoint_dir(tmp_path / "ckpt")
    assert set(sd2) == set(sd) and all(torch.equal(sd2[k].float(), sd[k]) for k in sd)
    gm = model.get_model()
    gm.initialize_vision_modules(gm.config)
    assert gm.get_vision_tower().to(dtype=torch.bfloat16, device="cpu") is not None
    assert model.bfloat16().cuda().eval() is model and model.module is model
    model.resize_token_embeddings(cfg.vocab_size)
    with pytest.raises(ValueError):
        model.resize_token_embeddings(cfg.vocab_size + 1)
    with pytest.raises(RuntimeError):
        model.float()


def _write_object_maps(tmp_path, n_verts=3000):
    import joblib
    op2v, obary = S.make_mesh_lift_maps(n_verts=n_verts, seed=3, coverage=0.25)
    pkl = tmp_path / "lift2d_dict.pkl"
    joblib.dump({"pixel_to_vertices_map": [op2v[v] for v in range(4)], "bary_coords_map": [obary[v] for v in range(4)],
                 "num_vertices": n_verts}, pkl)
    p2p = S.make_point_lift_maps(seed=2)
    mask_paths = []
    for v in range(4):
        np.savez(tmp_path / f"obj_p2pmap_{v}.npz", mapping=p2p[v])
        np.savez(tmp_path / f"obj_p2vmap_{v}.npz", pixel_to_vertices_map=op2v[v], bary_coords_map=obary[v], num_vertices=n_verts)
        mask_paths.append(str(tmp_path / f"obj_mask_{v}.png"))
    return pkl, mask_paths, (op2v, obary, n_verts), p2p


def test_object_paths_evaluate_and_forward(tmp_path, setup):
    """Rows a15 / a16 at model level: evaluate(contact_type='oafford', lift2d_dict_path=...) lifts through the object-mesh
    predictor (InteractVLM.py:624-628, sic precedence), model_forward with oC_loss_weight > 0 returns both object outputs
    with maps read from the files next to the mask paths (components.py:309, :363-375)."""
    from oracle import lift as OL
    cfg, sd, model, _ = setup
    pkl, mask_paths, (op2v, obary, nv), p2p = _write_object_maps(tmp_path)
    ids, ans, clip, sam, cam = tiny_inputs(cfg, 1)
    out = model.evaluate(clip, sam, ids, cam, [SIZE], [SIZE], lift2d_dict_path=str(pkl), contact_type="oafford",
                         max_new_tokens=ans.shape[1], scripted=ans)
    pm = torch.stack(out["pred_masks"], 0).cpu().numpy()
    want = OL.lift_object_mesh(pm, op2v, obary, nv, thr=0.3)
    assert out["pred_contact_3d"].shape == (1, nv)
    assert np.abs(out["pred_contact_3d"].cpu().numpy() - want).max() < 1e-5
    # teacher-forced path with object predictors switched on
    full = torch.cat([ids, ans], 1)
    model.oC_loss_weight, old = 3.0, model.oC_loss_weight
    try:
        for ds, key in (("piad_oafford", "pred_object_3d_afford"), ("pico_ocontact", "pred_object_3d_contact")):
            res = model(images=sam, images_clip=clip, input_ids=full, labels=full, attention_masks=torch.ones_like(full),
                        offset=torch.tensor([0, 1]), masks_list=[torch.zeros(4, 1, *SIZE)], label_list=[torch.zeros(SIZE)],
                        gt_contact_3d_list=[None], cam_params=cam, resize_list=[SIZE], ds_name_list=[ds],
                        mask_paths_list=[mask_paths], inference=True)
            pm = torch.stack(res["pred_masks"], 0).cpu().numpy()
            if "oafford" in ds:   # HM view type: the maps were sigmoid-ed before lifting (InteractVLM.py:452-456)
                assert pm.min() >= 0 and pm.max() <= 1
                assert np.abs(res[key].cpu().numpy() - OL.lift_points(pm, p2p, 2048)).max() < 1e-5
                assert res["pred_object_3d_contact"].shape == (1, 0)   # 'ocontact' not in ds_name (components.py:430)
            else:
                assert np.abs(res[key].cpu().numpy() - OL.lift_object_mesh(pm, op2v, obary, nv)).max() < 1e-5
                assert res["pred_object_3d_afford"].abs().max().item() == 0
            assert res["pred_human_3d_contact"].abs().max().item() == 0  # not an hcontact sample
    finally:
        model.oC_loss_weight = old


@pytest.mark.parametrize("cam_type,token_type,tok_name", [("simple", "Gen-Hu-Obj", "hseg"), ("view_index", "Gen-Int", "oseg"),
                                                          ("view_index", "Gen", "seg"), ("vi_v1", "Gen-Hu-Obj", "seg")])
def test_camera_encoder_and_token_type_variants(cam_type, token_type, tok_name):
    """The other camera encoders (components.py:491-539) and the AttentionSplitter token types (Gen-Hu-Obj / Gen-Int,
    InteractVLM.py:268-294,535-543): host logic with emulated kernels against the oracle (itself pinned to the reference's
    own classes, tests/test_oracle_components_golden.py), through the whole evaluate() call."""
    cfg = IVLMConfig.tiny()
    cfg.cam_encoder_type, cfg.token_type = cam_type, token_type
    cfg.hseg_token_idx, cfg.oseg_token_idx = 323, 324
    sd = S.make_state_dict(cfg, seed=5)
    for k in [k for k in sd if k.startswith(("attention_splitter.", "cam_pose_encoder.")) and k.endswith("weight")]:
        sd[k] = sd[k] * 1.5                      # a non-flat softmax (score spread ~2.4) without saturating it
    model = InteractVLMForCausalLM(cfg, sd, ctx=EmuContext())
    p2v, bary = S.make_mesh_lift_maps(seed=TINY_SEED["maps"])
    model.set_human_lift_maps(p2v, bary)
    ids, ans, clip, sam, cam = tiny_inputs(cfg, 1)
    ans = ans.clone()
    tok = {"seg": cfg.seg_token_idx, "hseg": cfg.hseg_token_idx, "oseg": cfg.oseg_token_idx}[tok_name]
    ans[ans == cfg.seg_token_idx] = tok
    out = model.evaluate(clip, sam, ids, cam, [SIZE], [SIZE], max_new_tokens=ans.shape[1], scripted=ans)
    st = {}
    ref = OM.evaluate(sd, cfg, clip, sam, ids, cam, [SIZE], [SIZE], lift_maps=(p2v, bary, S.N_SMPL),
                      max_new_tokens=ans.shape[1], scripted=ans, dtype=torch.float32, stages=st)
    assert torch.equal(out["output_ids"].cpu(), ref["output_ids"]) and out["pred_masks"][0].shape == (4, 1024, 1024)
    rm = ref["pred_masks"][0]
    assert (out["pred_masks"][0] - rm).abs().max().item() < 0.08 * rm.abs().max().item()
    assert (out["pred_contact_3d"] - ref["pred_contact_3d"]).abs().max().item() < 0.1   # bf16 storage vs the fp32 oracle
    # the prompt itself, stage-wise: bf16 storage only
    hid = OM.lm_hidden(OM.W(sd, torch.float32), cfg, clip.float(), ref["output_ids"][:, :-1])
    rows, toks = OM.seg_rows(cfg, ref["output_ids"], with_tokens=True)
    assert toks == [tok]
    prompt, _ = model.eng.seg_prompt(hid[0, rows[0]].bfloat16(), cam.bfloat16(), toks)
    want = st["prompt"][0]
    assert (prompt.float() - want).abs().max().item() < 0.03 * max(1.0, want.abs().max().item())
    if token_type != "Gen" and tok_name != "seg":   # the splitter branch changed the prompt
        plain, _ = model.eng.seg_prompt(hid[0, rows[0]].bfloat16(), cam.bfloat16(), [cfg.seg_token_idx])
        assert (plain.float() - prompt.float()).abs().max().item() > 0.05


def test_prompts_of_different_lengths_in_one_batch(setup):
    """generate() / evaluate() with ragged prompts (right-padded like the reference's collate_fn, per-sample positions and
    last-prompt rows) against the per-sample oracle run on the UNPADDED prompts (llava_arch.py:98-347 handles the padding in
    the reference; its evaluate() is batch 1)."""
    cfg, sd, model, (p2v, bary) = setup
    ids, ans, clip, sam, cam = tiny_inputs(cfg, 3)
    prompts = [ids[0], ids[1][:-3], ids[2][:-7]]                        # 22, 19 and 15 ids, each with its image token
    out = model.evaluate(clip, sam, prompts, cam, [SIZE] * 3, [SIZE] * 3, max_new_tokens=ans.shape[1], scripted=ans)
    assert out["pred_contact_3d"].shape == (3, S.N_SMPL)
    for b in range(3):
        ref = OM.evaluate(sd, cfg, clip[b:b + 1], sam[b:b + 1], prompts[b][None], cam[b:b + 1], [SIZE], [SIZE],
                          lift_maps=(p2v, bary, S.N_SMPL), max_new_tokens=ans.shape[1], scripted=ans[b:b + 1])
        n = prompts[b].numel() + ans.shape[1]
        assert out["output_ids"][b, :n].tolist() == ref["output_ids"][0].tolist()
        assert bool((out["output_ids"][b, n:] == cfg.pad_token_id).all())
        assert np.abs(out["pred_contact_3d"][b].numpy() - ref["pred_contact_3d"][0].numpy()).max() < 0.08
    # the padded-tensor form with explicit lengths gives the same thing, and so does greedy decoding
    L = max(p.numel() for p in prompts)
    padded = torch.full((3, L), cfg.pad_token_id, dtype=torch.int64)
    for b, p in enumerate(prompts):
        padded[b, : p.numel()] = p
    out2 = model.evaluate(clip, sam, padded, cam, [SIZE] * 3, [SIZE] * 3, max_new_tokens=ans.shape[1], scripted=ans)
    assert torch.equal(out2["output_ids"], out["output_ids"]) and torch.equal(out2["pred_contact_3d"], out["pred_contact_3d"])
    w = OM.W(sd)
    g_ids, _ = model.generate(clip, prompts, max_new_tokens=3)
    for b in range(3):
        with torch.no_grad():
            seq, _, _ = OM.greedy_generate(w, cfg, clip[b:b + 1], prompts[b][None], 3)
        n = seq.shape[1]
        assert g_ids[b, :n].tolist() == seq[0].tolist()


def test_separate_human_and_object_mask_decoders():
    """token_type '*-DifDe' (InteractVLM.py:44-53,114-122): 'hcontact' decodes with human_mask_decoder, 'ocontact' / 'oafford'
    with object_mask_decoder, anything else with the shared copy -- against the oracle, and the three decoders must differ."""
    cfg = IVLMConfig.tiny()
    cfg.token_type = "Gen-DifDe"
    sd = S.make_state_dict(cfg, seed=TINY_SEED["weights"])
    model = InteractVLMForCausalLM(cfg, sd, ctx=EmuContext())
    p2v, bary = S.make_mesh_lift_maps(seed=TINY_SEED["maps"])
    model.set_human_lift_maps(p2v, bary)
    ids, ans, clip, sam, cam = tiny_inputs(cfg, 1)
    outs = {}
    for ct in ("hcontact", "ocontact"):
        ev = model.evaluate(clip, sam, ids, cam, [SIZE], [SIZE], contact_type=ct, max_new_tokens=ans.shape[1], scripted=ans)
        ref = OM.evaluate(sd, cfg, clip, sam, ids, cam, [SIZE], [SIZE], contact_type=ct, max_new_tokens=ans.shape[1], scripted=ans,
                          dtype=torch.bfloat16)
        a, b = ev["pred_masks"][0].float(), ref["pred_masks"][0].float()
        assert (a - b).abs().max().item() < 0.06 * b.abs().max().item(), ct
        outs[ct] = a
    assert (outs["hcontact"] - outs["ocontact"]).abs().max().item() > 0.1 * outs["hcontact"].abs().max().item()
    full = torch.cat([ids, ans], 1)
    fw = model(images=sam, images_clip=clip, input_ids=full, labels=full, attention_masks=torch.ones_like(full), offset=torch.tensor([0, 1]),
               masks_list=[torch.zeros(4, 1, *SIZE)], label_list=[torch.zeros(SIZE)], gt_contact_3d_list=[None], cam_params=cam,
               resize_list=[SIZE], ds_name_list=["pico_ocontact"], mask_paths_list=[None], inference=True)
    assert (fw["pred_masks"][0].float() - outs["ocontact"]).abs().max().item() < 0.05 * outs["ocontact"].abs().max().item()
