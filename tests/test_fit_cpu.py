"""Host-side pieces of the pose-refinement driver (interactvlm_b200/fit.py) against goldens recorded from the reference's own
lines of optim/fit.py (oracle/make_goldens_fit.py), plus the file readers of optim/data_io.py on a synthetic sample folder."""
import json
from pathlib import Path

import numpy as np
import torch

from interactvlm_b200 import fit as FIT
from oracle.make_goldens_fit import CASES, inputs

GOLD = np.load(Path(__file__).parent / "golden" / "fit_init.npz")


def test_init_translation_and_normal_filter_match_the_reference_lines():
    for name, (seed, fc) in CASES.items():
        hv, hn, hc, ov, on, oc, mask, focal, pp = inputs(seed)
        t = torch.from_numpy
        cam = FIT.CameraParams(t(focal), t(pp))
        h_mask = t(hc) > 0.5
        tr = FIT.initial_translation(t(mask), t(hv), h_mask, cam)
        assert np.allclose(tr.numpy(), GOLD[f"{name}_translation"], rtol=0, atol=1e-6)
        o_mask, probs = FIT.filter_contacts_by_normals(t(hn), h_mask, t(on), t(oc), fc)
        assert np.array_equal(o_mask.numpy(), GOLD[f"{name}_o_mask"])
        assert np.array_equal(probs.numpy(), GOLD[f"{name}_o_probs"])
        if name == "one_sided":
            assert 0 < int(o_mask.sum()) < int((t(oc) > 0.3).sum())   # the filter removed some pairs and kept some


def test_camera_and_mesh_helpers():
    cam = FIT.get_camera_params(np.array([10.0, 20.0, 192.0, 256.0]), device="cpu")
    assert np.allclose(cam.focal_length.numpy(), [5000.0, 5000.0]) and np.allclose(cam.principal_point.numpy(), [106.0, 148.0])
    # unit cube: area-weighted centroid is the centre even with an uneven vertex distribution; outward normals
    v = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], np.float64)
    f = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4], [1, 2, 6], [1, 6, 5], [2, 3, 7], [2, 7, 6],
                  [3, 0, 4], [3, 4, 7]])
    assert np.allclose(FIT.mesh_centroid(v, f), [0.5, 0.5, 0.5])
    n = FIT.vertex_normals(torch.from_numpy(v).float(), torch.from_numpy(f))
    assert np.allclose(np.linalg.norm(n.numpy(), axis=1), 1.0, atol=1e-6)
    assert float((n * (torch.from_numpy(v).float() - 0.5)).sum(1).min()) > 0     # every normal points away from the centre


def test_load_params_reads_the_sample_folder(tmp_path):
    g = np.random.default_rng(0)
    hv = g.normal(size=(1, 50, 3)).astype(np.float32) + np.array([0, 0, 3], np.float32)
    hf = np.stack([np.arange(48), np.arange(48) + 1, np.arange(48) + 2], 1)
    np.savez(tmp_path / "osx_human2.npz", smpl_vertices=hv, smpl_faces=hf, bbox_2=np.array([[5.0, 6.0, 96.0, 128.0]]))
    np.savez(tmp_path / "hcontact_vertices.npz", pred_contact_3d_smplx=g.random(50).astype(np.float32), pred_contact_3d_smplh=g.random(40))
    np.savez(tmp_path / "ocontact_vertices.npz", pred_contact_3d=g.random((1, 4)).astype(np.float32))
    (tmp_path / "object_mesh.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nf 1 2 3\nf 1 3 4\nf 1 4 2\nf 2 4 3\n")
    m = np.zeros((32, 32), int)
    m[4:9, 7:20] = 1
    json.dump({"mask": m.tolist(), "bbox": [7, 4, 13, 5]}, open(tmp_path / "object_detection.json", "w"))
    json.dump({"mask": m.tolist()}, open(tmp_path / "human_detection.json", "w"))
    human, obj, cam = FIT.load_params(tmp_path / "osx_human2.npz", tmp_path / "object_mesh.obj", tmp_path / "object_detection.json",
                                      device="cpu")
    assert human.vertices.shape == (50, 3) and human.normals.shape == (50, 3) and human.contact_verts.shape == (50,)
    assert np.allclose((human.vertices + human.centroid_offset).numpy(), hv[0], atol=1e-5)     # centred on the area centroid
    assert obj.vertices.shape == (4, 3) and obj.contact_verts.shape == (4,) and obj.mask.shape == (32, 32) and int(obj.mask.sum()) == 65
    # centred, then y and z flipped (data_io.py:188-190)
    raw = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float64)
    c = FIT.mesh_centroid(raw, np.array([[0, 1, 2], [0, 2, 3], [0, 3, 1], [1, 3, 2]]))
    assert np.allclose(obj.vertices.numpy(), (raw - c) * np.array([1, -1, -1]), atol=1e-6)
    assert np.allclose(cam.focal_length.numpy(), [5000 / 192 * 96, 5000 / 256 * 128])
