"""oracle/model.py's camera conditioning and AttentionSplitter against outputs of the reference's own classes and
`process_embeddings` (tests/golden/components.npz, written by oracle/make_goldens_components.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from interactvlm_b200.config import IVLMConfig
from oracle import model as OM
from oracle.make_goldens_components import HSEG, OSEG, SEG, component_weights, inputs

GOLD = np.load(Path(__file__).parent / "golden" / "components.npz")


def variant(cam_type, token_type):
    cfg = IVLMConfig.tiny()
    cfg.cam_encoder_type, cfg.token_type, cfg.seg_token_idx, cfg.hseg_token_idx, cfg.oseg_token_idx = cam_type, token_type, SEG, HSEG, OSEG
    sd = {"cam_pose_encoder." + k: v for k, v in component_weights(cam_type).items()}
    sd.update({"attention_splitter." + k: v for k, v in component_weights("splitter").items()})
    return cfg, sd


@pytest.mark.parametrize("cam_type", ["simple", "view_index", "vi_v1"])
@pytest.mark.parametrize("token_type", ["Gen", "Gen-Hu-Obj"])
def test_process_embeddings_matches_reference(cam_type, token_type):
    cfg, sd = variant(cam_type, token_type)
    w = OM.W(sd, torch.float32)
    emb, cam = inputs()
    for name, tok in (("seg", SEG), ("hseg", HSEG), ("oseg", OSEG)):
        got = OM.process_embeddings(w, cfg, emb, cam, tok).numpy()
        want = GOLD[f"out/{cam_type}/{token_type}/{name}"]
        assert got.shape == want.shape == (1, 4, 256) and np.abs(got - want).max() < 2e-5 * max(1.0, np.abs(want).max())
    if token_type != "Gen":   # the two branches really differ, and differ from the plain embedding
        a, b, c = (GOLD[f"out/{cam_type}/{token_type}/{n}"] for n in ("hseg", "oseg", "seg"))
        assert np.abs(a - b).max() > 0.1 and np.abs(a - c).max() > 0.1


def test_seg_rows_with_human_and_object_tokens():
    cfg, _ = variant("vi_v1", "Gen-Hu-Obj")
    ids = torch.tensor([[1, 5, 6, HSEG, 7, 2], [1, 5, OSEG, 8, SEG, 2], [1, 5, 6, 7, 8, 2]])
    rows, toks = OM.seg_rows(cfg, ids, with_tokens=True)
    assert rows == [[3 - 1 + cfg.img_emb_len], [2 - 1 + cfg.img_emb_len, 4 - 1 + cfg.img_emb_len], []]
    assert toks == [HSEG, OSEG, None]
    cfg.token_type = "Gen"
    assert OM.seg_rows(cfg, ids) == [[], [4 - 1 + cfg.img_emb_len], []]
