"""GPU parity of the Render-Localise-Lift kernels (lift.cu) through the C ABI: against the numpy oracle on the
same seeded inputs, against the reference's golden outputs, and through size-independent properties."""
from pathlib import Path

import numpy as np
import pytest
import torch

from interactvlm_b200 import synthetic as S

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


@pytest.fixture(scope="module")
def data():
    from oracle.make_goldens import LIFT_SEED, OBJ_NVERTS
    return dict(seed=LIFT_SEED, obj_n=OBJ_NVERTS, gold=np.load(GOLD / "lift.npz"),
                logits=S.make_mask_logits(2, seed=LIFT_SEED["logits"]))


def test_human_lift_vs_reference_golden_and_oracle(ctx, data):
    from interactvlm_b200.ops import LIFT_HUMAN, LiftMap
    from oracle import lift as OL
    p2v, bary = S.make_mesh_lift_maps(seed=data["seed"]["maps"])
    m = LiftMap(ctx, p2v, bary, S.N_SMPL)
    out = m(torch.from_numpy(data["logits"]).cuda(), LIFT_HUMAN).cpu().numpy()
    ref = data["gold"]["human"]
    assert np.abs(out - ref).max() < 1e-6            # north_star: <= 1e-3; fp32 gather is far inside it
    assert np.array_equal(out >= 0.5, ref >= 0.5)    # bit-exact contact vertex set
    assert np.abs(out - OL.lift_human(data["logits"], p2v, bary, S.N_SMPL)).max() < 1e-6
    assert m.nnz == 3 * int(((p2v >= 0) & (p2v < S.N_SMPL)).all(-1).sum())


def test_object_mesh_and_points_lift(ctx, data):
    from interactvlm_b200.ops import LIFT_OBJECT_MESH, LIFT_POINTS, LiftMap
    p2v, bary = S.make_mesh_lift_maps(n_verts=data["obj_n"], seed=data["seed"]["obj"], coverage=0.25)
    m = LiftMap(ctx, p2v, bary, data["obj_n"])
    out = m(torch.from_numpy(data["logits"][:1]).cuda(), LIFT_OBJECT_MESH, 0.3).cpu().numpy()
    ref = data["gold"]["object_mesh"]
    # a pixel whose probability sits within 1 ulp of the 0.3 gate may flip between expf implementations
    assert np.abs(out - ref).max() < 1e-5
    p2p = S.make_point_lift_maps(seed=data["seed"]["points"])
    mp = LiftMap(ctx, p2p, None, 2048)
    heat = torch.sigmoid(torch.from_numpy(data["logits"])).cuda().contiguous()
    outp = mp(heat, LIFT_POINTS).cpu().numpy()
    assert np.abs(outp - data["gold"]["points"]).max() < 1e-5


def test_smplx_spmv(ctx, data):
    from interactvlm_b200.ops import CsrMatrix
    mapping = S.make_smplx_matrix(seed=0)
    csr = CsrMatrix(ctx, mapping)
    out = csr(torch.from_numpy(data["gold"]["human"]).cuda()).cpu().numpy()
    assert np.abs(out - data["gold"]["smplx"]).max() < 1e-6


def test_lift_properties_full_size(ctx):
    """Size-independent properties at the bench batch size: constant logits lift to a constant, batch entries are
    independent, output is invariant to the batch position."""
    from interactvlm_b200.ops import LIFT_HUMAN, LiftMap
    p2v, bary = S.make_mesh_lift_maps(seed=7)
    m = LiftMap(ctx, p2v, bary, S.N_SMPL)
    B = 8
    logits = torch.from_numpy(S.make_mask_logits(B, seed=8)).cuda()
    out = m(logits, LIFT_HUMAN)
    const = torch.full((1, 4, 1024, 1024), 1.25, device="cuda")
    oc = m(const, LIFT_HUMAN)
    seen = oc > 0
    p = 1 / (1 + np.exp(-1.25))
    assert (oc[seen] - p).abs().max().item() < 1e-6
    assert torch.equal(out[3:4], m(logits[3:4].contiguous(), LIFT_HUMAN))
    assert torch.equal(out.flip(0), m(logits.flip(0).contiguous(), LIFT_HUMAN))
    assert (out >= 0).all() and (out <= 1).all() and (out[:, ~seen[0]] == 0).all()


def test_lift_edge_cases(ctx):
    from interactvlm_b200.ops import LIFT_HUMAN, LiftMap
    p2v = np.full((2, 8, 8, 3), -1, np.int64)
    bary = np.zeros((2, 8, 8, 3), np.float32)
    p2v[0, 0, 0] = [0, 1, 2]; bary[0, 0, 0] = [0.5, 0.25, 0.25]
    p2v[0, 0, 1] = [0, 1, 99]; bary[0, 0, 1] = [0.3, 0.3, 0.4]
    m = LiftMap(ctx, p2v, bary, 4)
    assert m.nnz == 3
    masks = torch.zeros(1, 2, 8, 8, device="cuda")
    masks[0, 0, 0, 0] = 30.0
    out = m(masks, LIFT_HUMAN).cpu().numpy()
    assert np.allclose(out[0, :3], 1 / (1 + np.exp(-20.0)), atol=1e-7) and out[0, 3] == 0
    empty = LiftMap(ctx, np.full((1, 4, 4, 3), -1, np.int64), np.zeros((1, 4, 4, 3), np.float32), 5)
    assert empty.nnz == 0
    assert empty(torch.randn(2, 1, 4, 4, device="cuda"), LIFT_HUMAN).abs().max().item() == 0


def test_lowres_lift_is_bit_identical_to_upsample_then_lift(ctx, data):
    """ivlm_lift_lowres evaluates the x4 bilinear of Sam.postprocess_masks per map entry: same bits as ivlm_bilinear_f32
    followed by ivlm_lift, in all three modes, for a batch that spans two sample chunks (9 > LIFT_BC)."""
    from interactvlm_b200.ops import LIFT_HUMAN, LIFT_OBJECT_MESH, LIFT_POINTS, LiftMap
    from oracle import lift as OL
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    low = (torch.randn(9, 4, 256, 256, generator=g) * 4).bfloat16().float().cuda()   # bf16-valued like the decoder's output
    full = ctx.bilinear(low.view(36, 256, 256), 1024, 1024).view(9, 4, 1024, 1024)
    p2v, bary = S.make_mesh_lift_maps(seed=data["seed"]["maps"])
    m = LiftMap(ctx, p2v, bary, S.N_SMPL)
    a = m.lowres(low, LIFT_HUMAN)
    assert torch.equal(a, m(full, LIFT_HUMAN))
    assert torch.equal(m.lowres(low, LIFT_OBJECT_MESH, 0.3), m(full, LIFT_OBJECT_MESH, 0.3))
    # against torch's own interpolate + the numpy oracle (reference op order)
    ref_full = F.interpolate(low.cpu(), (1024, 1024), mode="bilinear", align_corners=False).numpy()
    assert np.abs(a.cpu().numpy() - OL.lift_human(ref_full, p2v, bary, S.N_SMPL)).max() < 1e-6
    p2p = S.make_point_lift_maps(seed=data["seed"]["points"])
    mp = LiftMap(ctx, p2p, None, 2048)
    assert torch.equal(mp.lowres(low, LIFT_POINTS), mp(full, LIFT_POINTS))


def test_lift_nonpositive_weight_sum_keeps_raw_votes(ctx):
    """components.py:257-262: votes are divided by the weight sum only where it is positive; elsewhere the raw votes are still
    added to the prediction (and the view is not counted)."""
    from interactvlm_b200.ops import LIFT_HUMAN, LIFT_OBJECT_MESH, LiftMap
    from oracle import lift as OL
    p2v = np.full((2, 4, 4, 3), -1, np.int64)
    bary = np.zeros((2, 4, 4, 3), np.float32)
    p2v[0, 1, 1] = [0, 1, 2]; bary[0, 1, 1] = [-0.25, 0.75, 0.5]     # vertex 0: negative weight sum in view 0
    p2v[1, 2, 2] = [0, 1, 3]; bary[1, 2, 2] = [0.5, 0.25, 0.25]      # and a regular vote in view 1
    m = LiftMap(ctx, p2v, bary, 5)
    logits = np.random.default_rng(0).normal(0, 2, (1, 2, 4, 4)).astype(np.float32)
    out = m(torch.from_numpy(logits).cuda(), LIFT_OBJECT_MESH, 0.0).cpu().numpy()
    ref = OL.lift_object_mesh(logits, p2v, bary, 5, thr=0.0)
    assert np.abs(out - ref).max() < 1e-6 and ref[0, 0] != 0
    outh = m(torch.from_numpy(logits).cuda(), LIFT_HUMAN).cpu().numpy()
    assert np.abs(outh - OL.lift_human(logits, p2v, bary, 5)).max() < 1e-6
