"""bench.py --workload joint_fit (BASELINE.json configs[3]): joint human + object contact -> pose refinement.

One step on a GPU = B samples: `evaluate(contact_type='hcontact')` on the B photos (4 body renders each) -> SMPL contact ->
SMPL-X (`convert_contacts` as CSR SpMV); `evaluate(contact_type='ocontact')` on the B objects (4 rendered views each, per-object
`lift2d_dict.pkl`) -> per-vertex object contact; then, per sample, the pose refinement of optim/fit.py (contact thresholds,
normal-filtered pairs, contact ICP, `max_iter` Adam iterations over soft silhouette + centroid + contact terms) on a synthetic
scene that carries those contact vectors' shapes.  Samples are independent: ranks shard them, one all-gather of the fitted
poses ([B, 9] = rot6d + translation).  The reference runs the three stages as separate scripts (run_demo.py twice, then
`python -m optim.fit`), one sample at a time."""
from __future__ import annotations

import json
import os
import tempfile

import numpy as np
import torch

from . import fit as FIT
from . import optim as PO
from . import synthetic as S


def _scene(size, seed, dev):
    """Human = large blob, object = small blob resting against it; target mask = the object's silhouette at the true pose."""
    hv, hf = S.make_test_mesh("blob", n_lat=24, n_lon=40, seed=seed)
    ov, of_ = S.make_test_mesh("blob", n_lat=14, n_lon=24, seed=seed + 1)
    hv = hv * 1.6 + np.array([0.0, 0.0, 3.0], np.float32)
    gt_t = np.array([0.62, 0.1, 2.75], np.float32)
    ov = ov * 0.5
    og = ov + gt_t
    d = np.linalg.norm(og[:, None, :] - hv[None, :, :], axis=-1)
    t = lambda a, dt=torch.float32: torch.as_tensor(np.asarray(a), dtype=dt, device=dev)
    cam = FIT.CameraParams(torch.tensor([1.1 * size, 1.1 * size], device=dev), torch.tensor([size / 2.0, size / 2.0], device=dev))
    ren = PO.SSRenderer((size, size), None, t(of_, torch.int64), {"focal_length": cam.focal_length, "principal_point": cam.principal_point},
                        device=dev)
    with torch.no_grad():
        mask = (ren.render(t(og), want_depth=False)[0][0, ..., 3] > 0.5).to(torch.uint8)
    hverts, overts = t(hv), t(ov)
    human = FIT.HumanParams(hverts, t(hf, torch.int64), FIT.vertex_normals(hverts, t(hf, torch.int64)),
                            t((d.min(0) < 0.06).astype(np.float32)), torch.zeros(3, device=dev))
    obj = FIT.ObjectParams(overts, t(of_, torch.int64), FIT.vertex_normals(overts, t(of_, torch.int64)),
                           t((d.min(1) < 0.06).astype(np.float32)), mask, torch.ones(1, device=dev))
    return human, obj, cam


def run(args, cfg, rank, local_rank, world, dist, model_name, ClockSampler, peaks):
    import bench as BM   # make_batch, N_ANS ... (bench.py is the entry script; importable from the repo root)
    from .harness import ContactConverter
    from .model import InteractVLMForCausalLM
    from .parallel import gather_contacts

    dev = torch.device("cuda", local_rank)
    batch, size, iters = args.batch, 256, int(os.environ.get("IVLM_FIT_ITERS", "250"))
    cfg.hC_loss_weight, cfg.oC_loss_weight = 3.0, 3.0
    sd = S.make_state_dict(cfg, seed=0, device=dev, gain=0.5)
    model = InteractVLMForCausalLM(cfg, sd, device=local_rank, use_pdl=bool(args.pdl))
    del sd
    p2v, bary = S.make_mesh_lift_maps(seed=0)
    model.set_human_lift_maps(p2v, bary)
    conv = ContactConverter(model, S.make_smplx_matrix(seed=0))
    ids, ans, clip_h, sam_h, cam_h = BM.make_batch(cfg, batch, seed=3234 + rank)
    _, _, oclip_h, osam_h, _ = BM.make_batch(cfg, batch, seed=4234 + rank)
    host = [t.pin_memory() for t in (clip_h, sam_h, cam_h, oclip_h, osam_h)]
    res = [t.to(dev) for t in host]
    sizes = [BM.SIZE] * batch
    tmp = tempfile.mkdtemp(prefix="ivlm_fit_")
    import joblib

    pkls = []
    for b in range(batch):   # one lift2d_dict.pkl per object (utils/demo_utils.py:171-257 writes them in the demo flow)
        nv = 2000 + 137 * b
        op2v, obary = S.make_mesh_lift_maps(n_verts=nv, seed=50 + b, coverage=0.25)
        pth = os.path.join(tmp, f"r{rank}_obj{b}_lift2d_dict.pkl")
        joblib.dump({"pixel_to_vertices_map": [op2v[v] for v in range(4)], "bary_coords_map": [obary[v] for v in range(4)],
                     "num_vertices": nv}, pth)
        pkls.append(pth)
    model.object_3d_contact_predictor.CACHE_ENTRIES = max(16, batch)
    scenes = [_scene(size, 10 * rank + b, dev) for b in range(batch)]
    opt = FIT.default_options()
    opt["max_iter"] = iters
    host_out = torch.empty((batch * world, 9), dtype=torch.float32).pin_memory()

    def step(resident, gather=True):
        c, s, k, oc, os_ = res if resident else [t.to(dev, non_blocking=True) for t in host]
        h = model.evaluate(c, s, ids, k, sizes, sizes, contact_type="hcontact", max_new_tokens=BM.N_ANS, scripted=ans)
        smplx = conv(h["pred_contact_3d"]).reshape(batch, -1)
        o = model.evaluate(oc, os_, ids, k, sizes, sizes, lift2d_dict_path=pkls, contact_type="ocontact", max_new_tokens=BM.N_ANS,
                           scripted=ans)
        assert smplx.shape[1] == S.N_SMPLX and len(o["pred_contact_3d"]) == batch
        fits = FIT.run_fit_many(scenes, (size, size), opt)   # the batch's fits advance together (interleaved graph replays)
        local = torch.stack([torch.cat([r.rotation6d.reshape(6), r.translation.reshape(3)]) for r in fits])
        allp = gather_contacts(local, dist) if gather else local
        if not resident:
            host_out[: allp.shape[0]].copy_(allp, non_blocking=True)
        return allp

    def timed(resident, steps):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        count = lambda: (model.launch_count() + PO._ctx(dev).launch_count() + PO.REPLAYED_LAUNCHES[0] +
                         sum(c.launch_count() for c in FIT._CTX_POOL.get(dev.index or 0, [])))
        n0 = count()
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
        return ms, count() - n0

    try:
        for _ in range(max(1, min(args.warmup, 3))):
            step(True)
        step(False)
        torch.cuda.synchronize()
        with ClockSampler(local_rank) as cs:
            ms, launches = timed(True, args.steps)
            ms_e2e, _ = timed(False, args.steps)
        # stage split of one step (CUDA events; host time of a stage shows up in the next event)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        c, s_, k, oc, os_ = res
        ev[0].record()
        model.evaluate(c, s_, ids, k, sizes, sizes, contact_type="hcontact", max_new_tokens=BM.N_ANS, scripted=ans)
        ev[1].record()
        model.evaluate(oc, os_, ids, k, sizes, sizes, lift2d_dict_path=pkls, contact_type="ocontact", max_new_tokens=BM.N_ANS, scripted=ans)
        ev[2].record()
        torch.cuda.synchronize()
        import time

        t0 = time.perf_counter()
        FIT.run_fit_many(scenes, (size, size), opt)
        torch.cuda.synchronize()
        fit_ms = (time.perf_counter() - t0) * 1e3 / batch   # wall clock: the fits run on their own streams
        stage = {"evaluate_hcontact": round(ev[0].elapsed_time(ev[1]), 1), "evaluate_ocontact": round(ev[1].elapsed_time(ev[2]), 1),
                 "fit_all_samples": round(fit_ms * batch, 1)}
    finally:
        import shutil

        shutil.rmtree(tmp, ignore_errors=True)
    images = batch * world * args.steps
    line = {"metric": "samples/sec (joint human+object contact -> pose refinement)", "value": images / (ms / 1e3), "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16 (contact inference) / f32 (fit)", "data": "synthetic",
            "config": {"workload": f"configs[3]: batch={batch}/GPU joint human+object contact ({model_name}; two evaluate() calls per "
                                   f"sample: hcontact with 4 body views, ocontact with 4 object views and a per-object lift2d_dict.pkl) -> "
                                   f"optim.fit pose refinement ({iters} Adam iterations at {size}^2, contact ICP init) per sample",
                       "batch_per_gpu": batch, "global_batch": batch * world, "fit_iterations": iters, "fit_image": size,
                       "parallelism": f"dp{world} (samples sharded, one NCCL all-gather of the fitted poses [B,9])"},
            "clocks": cs.summary(),
            "e2e": {"value": images / (ms_e2e / 1e3), "unit": "samples/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(sum(t.numel() * t.element_size() for t in host)) * world,
                    "d2h_bytes_per_step": int(host_out.numel() * 4)},
            "gpu_launches": int(launches), "stage_ms": stage, "fit_ms_per_sample": fit_ms, "fit_ms_per_iteration": fit_ms / max(iters, 1)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
