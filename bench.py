#!/usr/bin/env python
"""Throughput of the InteractVLM 3D human-contact hot path (BASELINE.json metric: images/sec, 1024x1024).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--workload hcontact|oafford_pc|joint_fit]
                    [--sweep 1,4,16,64,256] [--impl ours|reference]

Default workload = BASELINE.json configs[1] (`--workload hcontact`): one step = model.evaluate() on a batch of B = 8 synthetic
samples per GPU: one 224^2 CLIP image + V=4 SAM views of 1024^2 + a 74-id prompt -> greedy decode (scripted 24-token answer
through the paged KV cache) -> [SEG] prompt -> SAM ViT-H on the 4 views -> two-way mask decoder -> 4x1024^2 logits ->
per-vertex SMPL contact [6890]; LLaMA-2-13B + CLIP-L/14 + SAM ViT-H, random-init weights.
`--workload oafford_pc` = configs[2] (teacher-forced model_forward + 2048-point-cloud lift through p2pmap files),
`--workload joint_fit` = configs[3] (human + object contact -> optim.fit pose refinement), `--sweep` = configs[4].
`value` times the step with inputs resident in HBM; `e2e` times the same call with pinned-host inputs (H2D inside)
and the result read back (D2H inside).  N>1: one process per GPU (torchrun), the batch is sharded (B per rank, weak
scaling) and the per-sample result vectors are all-gathered over NCCL once per step.
`--impl reference` runs the CPU restatement of the reference algorithm (oracle/, batch 1, no KV cache, as the reference
executes it) on whole images on the host cores, wall clock, no extrapolation; the number of images is capped by a time budget.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from interactvlm_b200 import synthetic as S  # noqa: E402
from interactvlm_b200.config import IVLMConfig  # noqa: E402

METRIC = "images/sec (3D hcontact, 1024x1024)"
SIZE = (1024, 1024)
N_PRE, N_POST, N_ANS = 40, 30, 24  # prompt 1+40+3+30 = 74 ids (+255 image rows) and a 24-token scripted answer


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of this rank's GPU during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.rows.append([c.strip() for c in r.stdout.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "power_w_max": max(float(r[2]) for r in self.rows),
                "samples": len(self.rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------ workload
def make_batch(cfg, batch, seed):
    ids, ans = S.make_prompt_ids(cfg, batch, n_pre=N_PRE, n_post=N_POST, n_answer=N_ANS, seed=seed)
    rng = np.random.default_rng(seed + 6000)
    V = cfg.multiview_channels
    # uint8 U[0,255] pixels, normalised like run_demo.py:65-79 / CLIPImageProcessor, stored bf16 (the dtype the
    # reference hands to evaluate()); generated with torch to keep host RAM and time bounded at batch 8+
    g = torch.Generator().manual_seed(seed)
    clip = torch.randint(0, 256, (batch, 3, cfg.clip_image_size, cfg.clip_image_size), generator=g, dtype=torch.uint8)
    sam = torch.randint(0, 256, (batch, V, 3, cfg.sam_img_size, cfg.sam_img_size), generator=g, dtype=torch.uint8)
    cm = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)
    cs = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)
    sm = torch.tensor([123.675, 116.28, 103.53]).view(1, 1, 3, 1, 1)
    ss = torch.tensor([58.395, 57.12, 57.375]).view(1, 1, 3, 1, 1)
    clip = ((clip.float() / 255.0 - cm) / cs).bfloat16()
    sam = ((sam.float() - sm) / ss).bfloat16()
    cam = torch.from_numpy(np.broadcast_to(S.HCONTACT_CAM_PARAMS, (batch, V, 5)).copy()).bfloat16()
    del rng
    return torch.from_numpy(ids), torch.from_numpy(ans), clip, sam, cam


def algorithmic_flops(cfg, B, L, G):
    """2*MAC of the KV-cached path per batch, dense contractions + attention."""
    D, F, nl, V = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, cfg.multiview_channels
    S_ = L - 1 + cfg.clip_tokens - 1
    lin_tok = 2 * nl * (4 * D * D + 3 * D * F)
    llm = B * ((S_ + G - 1) * lin_tok + 2 * (1 + G - 1) * D * cfg.vocab_size)
    llm += B * nl * 4 * D * (S_ * S_ / 2 + sum(S_ + t for t in range(G - 1)))
    E, T = cfg.sam_embed_dim, cfg.sam_grid ** 2
    nwin = ((cfg.sam_grid + cfg.sam_window_size - 1) // cfg.sam_window_size) ** 2 * cfg.sam_window_size ** 2
    n_glob = len(cfg.sam_global_attn_indexes)
    n_win = cfg.sam_depth - n_glob
    per_view = 2 * T * 768 * E + cfg.sam_depth * 2 * T * 8 * E * E  # patch embed + MLP
    per_view += n_glob * (2 * T * 4 * E * E + 4 * T * T * E) + n_win * (2 * nwin * 4 * E * E + 4 * nwin * cfg.sam_window_size ** 2 * E)
    per_view += 2 * T * E * 256 + 2 * T * 2304 * 256
    C, CF, CT = cfg.clip_hidden_size, cfg.clip_intermediate_size, cfg.clip_tokens
    clip = cfg.clip_layers_used * (2 * CT * (4 * C * C + 2 * C * CF) + 4 * CT * CT * C) + 2 * 256 * 588 * C + 2 * 256 * C * D
    dec = 14.6e9 / 4  # per view
    return dict(llm=llm, sam=B * V * per_view, clip=B * clip, decoder=B * V * dec, total=llm + B * V * (per_view + dec) + B * clip)


# ------------------------------------------------------------------------------------------------ CPU reference arm
def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def fast_host_state_dict(cfg, dtype, gain=0.5, seed=1, only=None):
    """Random-init weights of the architecture on the HOST in `dtype`, built quickly: the big matrices are filled from a pool
    of 2^26 normal samples (scaled by the layer's init std, start offset varied per tensor) instead of 13e9 fresh draws.
    Real, distinct memory for every tensor -- the timing below is a memory-bound CPU run, so nothing may alias -- only
    the VALUES repeat, which timing does not see.  Small tensors go through the usual seeded generator."""
    spec = S.state_dict_spec(cfg)
    g = torch.Generator().manual_seed(seed)
    pool = torch.randn(1 << 26, generator=g).to(dtype)
    rng = np.random.default_rng(seed)
    out = {}
    for name, (shape, kind) in spec.items():
        if only is not None and not only(name):
            continue
        n = int(np.prod(shape))
        std, mean = 0.05, (1.0 if kind == "g" else 0.0)
        if kind == "w":
            std = gain / float(np.sqrt(S._fan_in(shape, name)))
            for suffix, gg in S._LOGIT_GAIN.items():
                if name.endswith(suffix):
                    std *= gg
        elif kind == "e":
            std = 0.5 if "embed_tokens" in name else 0.1
        elif kind == "pe":
            std = 1.0
        t = torch.empty(n, dtype=dtype)
        i = 0
        while i < n:
            off = int(rng.integers(0, 1 << 20))
            m = min(n - i, pool.numel() - off)
            torch.mul(pool[off:off + m], std, out=t[i:i + m])
            i += m
        if mean:
            t += mean
        out[name] = t.view(shape)
    return out


_CPU = {}


def cpu_reference(cfg, max_images=1, budget_s=0.0, threads=None, want_dtype=None):
    """Runs the oracle's restatement of the reference algorithm on the host cores the way the reference executes an image:
    `oracle.model.evaluate()` at batch 1 (the reference is batch-1 only) -- CLIP + all LLaMA layers over the
    WHOLE sequence for each of the N_ANS generated tokens (no KV cache), SAM ViT-H on 4 views, mask decoder,
    postprocess, lift -- whole images, wall clock, no extrapolation.  Runs images until `max_images` or until `budget_s`
    seconds have elapsed (at least one).  bf16 like the reference mandates (evaluate.py:532) unless fp32 is faster on this
    CPU (probed on one LLaMA layer; the faster one is used and named).
    Returns dict(value images/s, sec_per_image [..], cores, dtype, desc)."""
    from oracle import model as OM

    threads = threads or host_threads()
    torch.set_num_threads(threads)
    key = (cfg.hidden_size, cfg.num_hidden_layers, cfg.sam_depth)
    t_setup = time.perf_counter()
    if key not in _CPU:
        dtype = want_dtype
        if dtype is None:
            # probe: one LLaMA layer at the prompt length in both arithmetic types
            one = IVLMConfig.from_dict(cfg.to_dict())
            one.num_hidden_layers = 1
            S_ = N_PRE + N_POST + 4 - 1 + 256
            probe = {}
            for dt in (torch.bfloat16, torch.float32):
                sd1 = fast_host_state_dict(one, dt, only=lambda k: k.startswith("model.layers.0.") or k == "model.norm.weight")
                w1 = OM.W(sd1, dt)
                emb = torch.randn(1, S_, cfg.hidden_size).to(dt)
                with torch.no_grad():
                    OM.llama_forward(w1, one, emb)
                    t0 = time.perf_counter()
                    OM.llama_forward(w1, one, emb)
                    probe[dt] = time.perf_counter() - t0
            dtype = min(probe, key=probe.get)
            try:
                import psutil

                if dtype == torch.float32 and psutil.virtual_memory().available < 90e9 and cfg.hidden_size >= 4096:
                    dtype = torch.bfloat16
            except Exception:
                pass
        sd = fast_host_state_dict(cfg, dtype)
        p2v, bary = S.make_mesh_lift_maps(seed=0)
        _CPU[key] = (sd, dtype, (p2v, bary, S.N_SMPL))
    sd, dtype, maps = _CPU[key]
    t_setup = time.perf_counter() - t_setup
    secs = []
    t_all = time.perf_counter()
    for i in range(max(1, max_images)):
        ids, ans, clip, sam, cam = make_batch(cfg, 1, seed=4321 + i)
        t0 = time.perf_counter()
        out = OM.evaluate(sd, cfg, clip, sam, ids, cam, [SIZE], [SIZE], lift_maps=maps, max_new_tokens=N_ANS, scripted=ans,
                          dtype=dtype)
        secs.append(time.perf_counter() - t0)
        assert out["pred_contact_3d"] is not None and out["pred_contact_3d"].shape == (1, S.N_SMPL)
        if budget_s and time.perf_counter() - t_all + secs[-1] > budget_s:
            break
    sec = float(np.mean(secs))
    desc = (f"oracle.model.evaluate() = the reference algorithm as the reference runs it (batch 1, greedy generate WITHOUT KV cache: "
            f"{N_ANS} full passes of CLIP-L + {cfg.num_hidden_layers} LLaMA layers over {N_PRE + N_POST + 3 + 256}..+{N_ANS - 1} positions, "
            f"SAM ViT {cfg.sam_depth} blocks on 4 views of 1024^2, mask decoder, postprocess, lift), {len(secs)} whole image(s) "
            f"of the bench workload timed by wall clock on {threads} host threads in {str(dtype).replace('torch.', '')}: "
            f"{', '.join(f'{x:.1f}' for x in secs)} s per image; not extrapolated (weights built in {t_setup:.0f} s, untimed)")
    return dict(value=1.0 / sec, sec_per_image=secs, cores=threads, dtype=dtype, desc=desc)


# ------------------------------------------------------------------------------------------------ workloads
class HContact:
    """BASELINE.json configs[1]: model.evaluate() on B samples -> [B,6890] SMPL contact."""
    name, n_out = "hcontact", S.N_SMPL

    def __init__(self, cfg, model, batch, rank, dev, tmp=None):
        self.cfg, self.model, self.batch, self.dev = cfg, model, batch, dev
        p2v, bary = S.make_mesh_lift_maps(seed=0)
        model.set_human_lift_maps(p2v, bary)
        self.ids, self.ans, clip_h, sam_h, cam_h = make_batch(cfg, batch, seed=1234 + rank)
        self.host = [t.pin_memory() for t in (clip_h, sam_h, cam_h)]
        self.res = [t.to(dev) for t in self.host]
        self.sizes = [SIZE] * batch

    def describe(self):
        return (f"configs[1]: batch={self.batch}/GPU synthetic RGB, 3D human-contact (DAMON-shape), prompt {N_PRE + N_POST + 4} ids "
                f"(+255 image rows), {N_ANS} scripted answer tokens through the paged KV cache, V=4 views 1024^2 -> SMPL contact [6890]")

    def h2d_bytes(self):
        return int(sum(t.numel() * t.element_size() for t in self.host))

    def step(self, resident):
        c, s, k = self.res if resident else [t.to(self.dev, non_blocking=True) for t in self.host]
        out = self.model.evaluate(c, s, self.ids, k, self.sizes, self.sizes, contact_type="hcontact", max_new_tokens=N_ANS,
                                  scripted=self.ans)
        return out["pred_contact_3d"]

    def flops(self):
        return algorithmic_flops(self.cfg, self.batch, self.ids.shape[1], N_ANS)


class OAffordPC:
    """BASELINE.json configs[2]: object affordance on 2048-point clouds (LEMON/PIAD shape).  model(**input_dict) with
    inference=True -- ONE teacher-forced causal pass over prompt + answer (evaluate.py:120 `inference_type=forward`,
    InteractVLM.py:296-474) -- then ObjectPCAfford3DPredictor through per-sample `p2pmap_*.npz` files
    (components.py:289-347, heat-map view type 4MV-Z_HM: sigmoid-ed maps are lifted, InteractVLM.py:452-456)."""
    name, n_out = "oafford_pc", 2048

    def __init__(self, cfg, model, batch, rank, dev, tmp):
        self.cfg, self.model, self.batch, self.dev = cfg, model, batch, dev
        ids, ans, clip_h, sam_h, cam_h = make_batch(cfg, batch, seed=2234 + rank)
        self.ids = torch.cat([ids, ans], 1)                       # teacher forcing: the answer (with its [SEG]) is part of the input
        self.host = [t.pin_memory() for t in (clip_h, sam_h, cam_h)]
        self.res = [t.to(dev) for t in self.host]
        self.sizes = [SIZE] * batch
        V = cfg.multiview_channels
        self.mask_paths = []
        for b in range(batch):                                    # the files the dataset would point at
            p2p = S.make_point_lift_maps(seed=100 * rank + b)
            paths = []
            for v in range(V):
                mp = os.path.join(tmp, f"r{rank}_obj{b}_mask_{v}.png")
                np.savez(mp.replace("mask", "p2pmap")[:-4] + ".npz", mapping=p2p[v])
                paths.append(mp)
            self.mask_paths.append(paths)
        model.object_3d_afford_predictor.CACHE_ENTRIES = max(16, batch)
        self.labels = [torch.zeros(SIZE) for _ in range(batch)]   # only their shape is read (original size of the masks)

    def describe(self):
        return (f"configs[2]: batch={self.batch}/GPU synthetic RGB, object affordance (LEMON/PIAD-shape 2048-point clouds), teacher-forced "
                f"model_forward(inference=True) over {self.ids.shape[1]} ids (+255 image rows), V=4 views 1024^2, per-sample p2pmap files -> [2048]")

    def h2d_bytes(self):
        return int(sum(t.numel() * t.element_size() for t in self.host))

    def step(self, resident):
        c, s, k = self.res if resident else [t.to(self.dev, non_blocking=True) for t in self.host]
        out = self.model(images=s, images_clip=c, input_ids=self.ids, cam_params=k, resize_list=self.sizes, label_list=self.labels,
                         ds_name_list=["oafford"] * self.batch, mask_paths_list=self.mask_paths, inference=True)
        return out["pred_object_3d_afford"]

    def flops(self):
        f = algorithmic_flops(self.cfg, self.batch, self.ids.shape[1] , 1)
        return f


WORKLOADS = {"hcontact": HContact, "oafford_pc": OAffordPC}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8, help="samples per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="full", choices=["full", "tiny"])
    ap.add_argument("--workload", default="hcontact", choices=sorted(WORKLOADS) + ["joint_fit"],
                    help="hcontact = BASELINE configs[1] (the headline); oafford_pc = configs[2]; joint_fit = configs[3]")
    ap.add_argument("--sweep", default="", help="comma-separated batch sizes per GPU (BASELINE configs[4]: 1,4,16,64,256): one JSON line each")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="--impl reference: stop starting new whole images after this many seconds")
    ap.add_argument("--overlap", type=int, default=0, help="1: SAM encoder on a second stream next to the decode steps "
                    "(measured slower on B200, profiles/r1_overlap_timeline.txt; kept as an option)")
    ap.add_argument("--sm-limit", type=int, default=104, help="SMs the encoder GEMMs keep to while decode steps are in flight")
    ap.add_argument("--limited-chunks", type=int, default=-1, help="encoder chunks launched SM-limited (-1: estimate)")
    ap.add_argument("--sam-chunk", type=int, default=0, help="views per encoder chunk (0: model default, 16; 4 with --overlap)")
    ap.add_argument("--no-view-cache-pass", action="store_true", help="skip the separately reported constant-views measurement")
    ap.add_argument("--pdl", type=int, default=1, help="1: programmatic dependent launch for the decode chain")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = IVLMConfig.full() if args.config == "full" else IVLMConfig.tiny()
    if args.workload == "oafford_pc":      # interactvlm-3d-oafford-lemon-piad (scripts/run_train.sh): object loss on, heat-map views
        cfg.hC_loss_weight, cfg.oC_loss_weight, cfg.oC_sam_view_type = 0.0, 3.0, "4MV-Z_HM"
    model_name = "LLaMA-2-13B + CLIP-L/14 + SAM ViT-H" if args.config == "full" else "TINY debug config"

    if args.impl == "reference":
        if rank != 0:
            return
        return reference_arm(args, cfg, model_name)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (sm_100a); there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from interactvlm_b200 import build as _b

    if _b.needs_build():
        _b.build()
    if args.workload == "joint_fit":
        from interactvlm_b200 import bench_fit

        return bench_fit.run(args, cfg, rank, local_rank, world, dist, model_name, ClockSampler, peaks)
    from interactvlm_b200.model import InteractVLMForCausalLM
    from interactvlm_b200.parallel import gather_contacts

    dev = torch.device("cuda", local_rank)
    sd = S.make_state_dict(cfg, seed=0, device=dev, gain=0.5)
    model = InteractVLMForCausalLM(cfg, sd, device=local_rank, use_pdl=bool(args.pdl))
    del sd
    if args.overlap:
        model.enable_overlap(sm_limit=args.sm_limit, limited_chunks=None if args.limited_chunks < 0 else args.limited_chunks,
                             sam_chunk=args.sam_chunk or 4)
    elif args.sam_chunk:
        model.sam_chunk = args.sam_chunk
    import tempfile

    tmp = tempfile.mkdtemp(prefix="ivlm_bench_")
    batches = [int(x) for x in args.sweep.split(",") if x] or [args.batch]
    try:
        for bi, batch in enumerate(batches):
            line = run_workload(args, cfg, model, batch, rank, local_rank, world, dist, dev, tmp, model_name, gather_contacts,
                                extras=(bi == len(batches) - 1 and not args.sweep))
            if rank == 0:
                print(json.dumps(line), flush=True)
    finally:
        import shutil

        shutil.rmtree(tmp, ignore_errors=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def reference_arm(args, cfg, model_name):
    """`--impl reference`: the reference's own CPU path (its algorithm as restated by oracle/, kind "port": the reference tree
    and its pinned stack do not exist on the GPU box) on this box's host cores.  One step = one WHOLE image (batch 1, the only
    batch size the reference supports); warm-up = one tiny-config image (thread pools, oneDNN) instead of W whole images and
    the number of timed images is capped by --ref-budget-s so that the run ends within a few minutes -- both stated in the line."""
    t_start = time.perf_counter()
    if args.config == "full":
        tiny = IVLMConfig.tiny()
        cpu_reference(tiny, max_images=1, want_dtype=torch.float32)
    r = cpu_reference(cfg, max_images=max(1, args.steps), budget_s=args.ref_budget_s)
    sec = float(np.mean(r["sec_per_image"]))
    n = len(r["sec_per_image"])
    config = {"workload": f"configs[1] hcontact, {model_name}: the same per-image work as the product arm (prompt {N_PRE + N_POST + 4} ids + 255 "
                          f"image rows, {N_ANS} answer tokens, V=4 views 1024^2), executed image by image (batch 1) like the reference does",
              "images_per_step": 1, "global_batch": 1, "views": cfg.multiview_channels, "parallelism": "host cores only",
              "requested": {"steps": args.steps, "warmup": args.warmup},
              "cap": f"{n} of {args.steps} requested steps ran (no new image is started after {args.ref_budget_s:.0f} s); warm-up is one "
                     f"tiny-config image, not {args.warmup} whole images -- a whole image costs ~{sec:.0f} s on this box"}
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "images/s", "n_gpus": args.gpus, "steps": n,
            "warmup": 0, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": str(r["dtype"]).replace("torch.", "").replace("bfloat16", "bf16").replace("float32", "f32"), "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": "port", "sample": r["desc"]},
            "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "wall_s": time.perf_counter() - t_start}
    print(json.dumps(line))


def run_workload(args, cfg, model, batch, rank, local_rank, world, dist, dev, tmp, model_name, gather_contacts, extras=True):
    wl = WORKLOADS[args.workload](cfg, model, batch, rank, dev, tmp)
    config = {"workload": wl.describe() + f", {model_name}", "batch_per_gpu": batch, "global_batch": batch * world,
              "views": cfg.multiview_channels, "parallelism": f"dp{world} (batch-sharded, one NCCL all-gather of [B,{wl.n_out}])",
              "l2": f"inputs ({wl.h2d_bytes() / 1e6:.0f} MB/step) and weights (28 GB) exceed the 126 MB L2; no explicit flush",
              "overlap": ({"sam_encoder_next_to_decode": True, "sm_limit": args.sm_limit, "limited_chunks": args.limited_chunks,
                           "sam_chunk": args.sam_chunk} if args.overlap else None),
              "pdl_decode_chain": bool(args.pdl)}
    host_out = torch.empty((batch * world, wl.n_out), dtype=torch.float32).pin_memory()

    def step(resident: bool, gather=True):
        local = wl.step(resident)
        allc = gather_contacts(local, dist) if gather else local   # the one collective of the path (evaluate.py:185-222)
        if not resident:
            host_out[: allc.shape[0]].copy_(allc, non_blocking=True)
        return allc

    def timed(resident, steps):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = model.launch_count()
        e0.record()
        for _ in range(steps):
            step(resident)
        e1.record()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, model.launch_count() - n0

    for _ in range(args.warmup):
        step(True)
    step(False)
    torch.cuda.synchronize()
    with ClockSampler(local_rank) as cs:
        ms, launches = timed(True, args.steps)
        ms_e2e, _ = timed(False, args.steps)
    clocks = cs.summary()
    # ---- everything below runs on EVERY rank without any collective (so that no rank waits in a barrier while another
    # ---- one issues an all-gather); rank 0 reports
    model.record_stages = True
    step(True, gather=False)
    stages = model.stage_ms()
    model.record_stages = False
    # launches inside CUDA-graph replays are not seen by the handle's counter: add them explicitly
    graph_launches = 0
    st = next(iter(model._graphs.values()), None)
    if args.workload == "hcontact" and st is not None and st.get("graph_scripted") is not None:
        graph_launches = st.get("graph_launches", 0) * (N_ANS - 1) * args.steps
    images = batch * world * args.steps
    value = images / (ms / 1e3)
    e2e = images / (ms_e2e / 1e3)
    line = {"metric": METRIC if args.workload == "hcontact" else f"images/sec ({args.workload}, 1024x1024)", "value": value,
            "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e, "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": wl.h2d_bytes() * world, "d2h_bytes_per_step": int(host_out.numel() * 4)},
            "gpu_launches": int(launches + graph_launches), "sam_views_per_s": value * cfg.multiview_channels,
            "stage_ms": {k: round(v, 2) for k, v in stages.items()}}
    pk = peaks()
    fl = wl.flops()
    line["algorithmic_tflop_per_step_per_gpu"] = fl["total"] / 1e12
    line["path_tensor_frac"] = (fl["total"] / (ms / args.steps / 1e3) / 1e12) / pk["tf_sustained"]
    if "llm_decode" in stages and args.workload == "hcontact":
        # decode steps stream every LLaMA weight once per step for the whole batch: HBM-bound
        nl, D, F = cfg.num_hidden_layers, cfg.hidden_size, cfg.intermediate_size
        wbytes = 2.0 * (nl * (4 * D * D + 3 * D * F) + D * cfg.vocab_size)
        kv = 2.0 * 2 * nl * D * batch * (wl.ids.shape[1] + 255 + N_ANS / 2)
        per_step = stages["llm_decode"] / (N_ANS - 1) / 1e3
        line["decode_hbm"] = {"bound": "hbm", "achieved": (wbytes + kv) / per_step / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                              "frac": (wbytes + kv) / per_step / 1e9 / pk["hbm"], "ms_per_token_step": per_step * 1e3}

    if extras and world == 1 and not args.no_view_cache_pass and args.workload == "hcontact":
        # NOT the headline: the real hcontact harness feeds the SAME four body renders with every image (run_demo.py:279-281);
        # with the exact-match view cache the encoder runs once per distinct view instead of once per sample.
        model.enable_view_cache()
        clip_d, sam_d, cam_d = wl.res
        same = sam_d[:1].expand(batch, *sam_d.shape[1:]).contiguous()

        def cstep():
            out = model.evaluate(clip_d, same, wl.ids, cam_d, wl.sizes, wl.sizes, contact_type="hcontact", max_new_tokens=N_ANS,
                                 scripted=wl.ans)
            return out["pred_contact_3d"]

        for _ in range(2):
            cstep()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            cstep()
        e1.record()
        torch.cuda.synchronize()
        cms = e0.elapsed_time(e1) / args.steps
        line["hcontact_constant_views"] = {
            "value": batch / (cms / 1e3), "unit": "images/s", "ms_per_step": cms,
            "cache": {k: model._view_cache[k] for k in ("hits", "misses")},
            "note": "separate from the headline: all samples share the same 4 SAM views (as run_demo.py hcontact does), exact-match "
                    "view cache on (model.enable_view_cache): bit-identical outputs, encoder skipped for views seen before"}
        model._view_cache = None
        del same

    if not args.no_profile:
        # roofline of the dominant kernel (tcgen05 GEMM), timed per launch with CUDA events on the launching stream; every rank
        # runs the pass on its own shard, without the all-gather
        model.use_cuda_graph = False
        model._graphs = {}
        model.ctx.set_option("pdl", 0)  # per-launch event pairs need serialised kernels
        model.ctx.enable_profile()
        # park the GPU (~100 ms spin) at the start of the CLIP/prefill stage and of every SAM chunk so that the host
        # is hundreds of launches ahead: the event pairs then bracket pure GPU time, not host launch latency
        model.stage_delay = lambda: torch.cuda._sleep(200_000_000)
        step(True, gather=False)
        model.stage_delay = None
        rep = model.ctx.profile_report()
        model.ctx.disable_profile()
        model.ctx.set_option("pdl", int(bool(args.pdl)))
        model.use_cuda_graph = True
        tot = sum(r["ms"] for r in rep.values())
        g = rep["gemm"]
        ach = g["work"] / (g["ms"] / 1e3) / 1e12
        traffic, traffic_note = None, None
        tp = ROOT / "profiles" / "r2_traffic.json"
        if tp.exists():  # DRAM bytes of one representative launch from the committed ncu --set full capture (tools/ncu_traffic.py)
            t = json.loads(tp.read_text())
            k0 = t.get("representative") or next(iter(t["launches"]))
            traffic = t["launches"][k0]["dram_bytes"]
            traffic_note = {"launch": k0, "algorithmic_bytes_per_launch": t["launches"][k0]["algorithmic_bytes"],
                            "source": "profiles/r2_traffic.json (ncu --set full of the shipping kernel, dram__bytes_read.sum + "
                                      "dram__bytes_write.sum; regenerate with tools/ncu_traffic.py)"}
        line["roofline"] = {"kernel": "gemm_bf16_tcgen05_kernel (token count > 64: SAM, CLIP, LLaMA prefill)", "bound": "tensor",
                            "achieved": ach, "peak": pk["tf_sustained"],
                            "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"], "traffic": traffic,
                            "peak_source": pk["source"] + " sustained cuBLAS bf16 (kernel timed inside a long step)",
                            "traffic_note": traffic_note,
                            "launches_per_step": g["launches"], "share_of_kernel_time": g["ms"] / tot}
        line["kernel_time_shares"] = {k: round(r["ms"] / tot, 4) for k, r in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])}
        line["kernel_ms_eager_pass"] = {k: round(r["ms"], 2) for k, r in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])}
        if "sam_attention" in rep:
            a = rep["sam_attention"]
            line["sam_attention_tflops"] = a["work"] / (a["ms"] / 1e3) / 1e12
        line["note_small_m_gemm"] = ("decode-step GEMMs (token count <= 64, swapped operands) are weight streaming and are "
                                     "reported under decode_hbm from the graph-replayed decode stage; in this eager profiling "
                                     "pass their event times include host launch latency")
    if rank == 0 and extras and not args.no_cpu_baseline and world == 1 and args.workload == "hcontact":
        # same function as `--impl reference`, one whole image (no extrapolation)
        if args.config == "full":
            cpu_reference(IVLMConfig.tiny(), max_images=1, want_dtype=torch.float32)
        r = cpu_reference(cfg, max_images=1)
        line["cpu_baseline"] = {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": "port", "sample": r["desc"]}
    del wl
    return line


if __name__ == "__main__":
    main()
