"""GPU: input-preparation kernel (bit-exact vs the reference's preprocess formula in fp32 -> bf16), SMPL->SMPL-X SpMV and the
batched validate loop through the CUDA path."""
import numpy as np
import pytest
import torch

from interactvlm_b200 import harness as Hn
from interactvlm_b200 import synthetic as S
from interactvlm_b200.config import IVLMConfig
from oracle import lift as OL
from oracle.make_goldens_model import TINY_SEED, tiny_inputs

pytestmark = pytest.mark.gpu
SIZE = (1024, 1024)


@pytest.fixture(scope="module")
def model(ctx):
    from interactvlm_b200.model import InteractVLMForCausalLM

    cfg = IVLMConfig.tiny()
    sd = S.make_state_dict(cfg, seed=TINY_SEED["weights"])
    m = InteractVLMForCausalLM(cfg, sd, ctx=ctx)
    p2v, bary = S.make_mesh_lift_maps(seed=TINY_SEED["maps"])
    m.set_human_lift_maps(p2v, bary)
    return m


def test_preprocess_kernel_bit_exact(model):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (2, 224, 224, 3), dtype=np.uint8)
    views = rng.integers(0, 256, (2, 4, 683, 1024, 3), dtype=np.uint8)   # padded at the bottom
    clip, sam, resize = Hn.prepare_inputs(model, img, views)
    assert resize == [(683, 1024)] * 2
    x = torch.from_numpy(views).permute(0, 1, 4, 2, 3).float().cuda()
    ref = (x - torch.tensor([123.675, 116.28, 103.53], device="cuda").view(1, 1, 3, 1, 1)) / \
        torch.tensor([58.395, 57.12, 57.375], device="cuda").view(1, 1, 3, 1, 1)
    ref = torch.nn.functional.pad(ref, (0, 0, 0, 1024 - 683)).bfloat16()
    assert torch.equal(sam, ref)
    c = torch.from_numpy(img).permute(0, 3, 1, 2).float().cuda() * (1.0 / 255.0)
    cref = ((c - torch.tensor([0.48145466, 0.4578275, 0.40821073], device="cuda").view(1, 3, 1, 1))
            / torch.tensor([0.26862954, 0.26130258, 0.27577711], device="cuda").view(1, 3, 1, 1)).bfloat16()
    assert (clip.float() - cref.float()).abs().max().item() <= 2 ** -6  # <= 1 bf16 ulp at |x| < 4 (x/255 vs x*(1/255))


def test_convert_contacts_spmv_and_validate(model):
    mapping = S.make_smplx_matrix(seed=0)
    conv = Hn.ContactConverter(model, mapping)
    cfg = model.config
    ids, ans, clip, sam, cam = tiny_inputs(cfg, 2)
    ref = model.evaluate(clip, sam, ids, cam, [SIZE] * 2, [SIZE] * 2, max_new_tokens=ans.shape[1], scripted=ans)
    smplx = conv(ref["pred_contact_3d"])
    want = OL.convert_contacts(ref["pred_contact_3d"].cpu().numpy(), mapping)
    assert smplx.shape == (2, S.N_SMPLX) and np.abs(smplx.cpu().numpy() - want).max() < 1e-5
    samples = [dict(images_clip=clip[b], images=sam[b], input_ids=ids[b], cam_params=cam[b], resize=SIZE, original_size=SIZE,
                    scripted=ans[b], gt_contact_3d=(ref["pred_contact_3d"][b] >= 0.5).float()) for b in range(2)]
    preds, metrics = Hn.validate(model, samples, batch_size=2, max_new_tokens=ans.shape[1])
    assert torch.equal(preds, ref["pred_contact_3d"]) and metrics["f1"] > 0.999


@pytest.mark.parametrize("h,w,oh,ow,filt", [(480, 640, 768, 1024, "bilinear"), (1365, 2048, 683, 1024, "bilinear"),
                                           (480, 640, 224, 298, "bicubic"), (100, 37, 224, 82, "bicubic")])
def test_resize_kernel_bit_exact_vs_pillow(model, h, w, oh, ow, filt):
    from PIL import Image

    rng = np.random.default_rng(h + w)
    img = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
    got = model.ctx.resize_u8(torch.from_numpy(img).cuda(), oh, ow, filt).cpu().numpy()
    for i in range(2):
        want = np.asarray(Image.fromarray(img[i]).resize((ow, oh), Image.BILINEAR if filt == "bilinear" else Image.BICUBIC))
        assert np.array_equal(got[i], want)


def test_prepare_inputs_from_raw_matches_reference_pipeline(model):
    """Raw uint8 images -> GPU resize + normalise == the reference's CPU pipeline (Pillow resize, preprocess, bf16 cast)."""
    from PIL import Image

    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (1, 480, 640, 3), dtype=np.uint8)
    views = rng.integers(0, 256, (1, 4, 600, 800, 3), dtype=np.uint8)
    clip, sam, resize = Hn.prepare_inputs_from_raw(model, img, views)
    assert resize == [(768, 1024)] and sam.shape == (1, 4, 3, 1024, 1024) and clip.shape == (1, 3, 224, 224)
    v = np.array(Image.fromarray(views[0, 1]).resize((1024, 768), Image.BILINEAR))
    x = torch.from_numpy(v).permute(2, 0, 1).float()
    ref = (x - torch.tensor([123.675, 116.28, 103.53]).view(-1, 1, 1)) / torch.tensor([58.395, 57.12, 57.375]).view(-1, 1, 1)
    ref = torch.nn.functional.pad(ref, (0, 0, 0, 1024 - 768)).bfloat16()
    assert torch.equal(sam[0, 1].cpu(), ref)
    c = np.asarray(Image.fromarray(img[0]).resize((298, 224), Image.BICUBIC))[:, 37:37 + 224]
    cref = ((torch.from_numpy(c.copy()).permute(2, 0, 1).float() * (1 / 255.0) - torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(-1, 1, 1))
            / torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(-1, 1, 1)).bfloat16()
    assert (clip[0].float().cpu() - cref.float()).abs().max().item() <= 2 ** -6


def test_jpeg_decode_on_the_gpu_matches_opencv_within_decoder_tolerance(ctx):
    """ivlm_jpeg_decode_rgb (nvJPEG) against cv2.imdecode (libjpeg-turbo, what the reference's cv2.imread runs): same size,
    pixels within a few grey levels (the two IDCT / chroma-upsampling implementations differ), 4:2:0 and 4:4:4, odd sizes."""
    import cv2

    g = np.random.default_rng(0)
    for (H, W), quality, sub in (((357, 500), 95, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420), ((224, 224), 90, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444)):
        yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
        img = np.stack([127 + 100 * np.sin(xx / 37 + c) * np.cos(yy / 29 - c) for c in range(3)], -1)
        img = np.clip(img + g.normal(0, 3, img.shape), 0, 255).astype(np.uint8)
        ok, enc = cv2.imencode(".jpg", img[..., ::-1], [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sub])
        assert ok
        ref = cv2.cvtColor(cv2.imdecode(enc, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
        out = ctx.decode_jpeg(enc.tobytes()).cpu().numpy()
        assert out.shape == ref.shape == (H, W, 3)
        d = np.abs(out.astype(np.int32) - ref.astype(np.int32))
        print(f"jpeg {H}x{W} q{quality}: max |diff| {d.max()}, mean {d.mean():.3f}")
        # 4:4:4: only the IDCT / colour-conversion rounding differs; 4:2:0: libjpeg-turbo smooths the upsampled chroma
        # ("fancy upsampling") and nvJPEG does not, which moves many pixels by one or two levels
        if sub == cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444:
            assert d.max() <= 4 and d.mean() < 0.8, (d.max(), d.mean())
        else:
            assert d.max() <= 24 and d.mean() < 2.0, (d.max(), d.mean())
    with pytest.raises(RuntimeError):
        ctx.decode_jpeg(b"not a jpeg at all")
