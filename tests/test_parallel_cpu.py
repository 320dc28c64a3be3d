"""CPU, world_size 2, gloo: the batch-shard + all-gather of per-sample contact vectors (interactvlm_b200/parallel.py),
which is the only exchange on the path (replaces evaluate.py:185-222 of the reference)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from interactvlm_b200.parallel import gather_contacts, pad_shard, shard_range


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 9, 64, 1370):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) <= (n + world - 1) // world


def test_pad_shard():
    x = torch.arange(6.0).view(2, 3)
    y = pad_shard(x, 4)
    assert y.shape == (4, 3) and torch.equal(y[:2], x) and y[2:].abs().sum() == 0
    assert pad_shard(x, 2) is not None and pad_shard(x, 2).shape == (2, 3)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_samples, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_samples * 5, dtype=torch.float32).view(n_samples, 5) + 0.25
        lo, hi = shard_range(n_samples, rank, world)
        got = gather_contacts(full[lo:hi].clone(), dist, n_samples=n_samples)
        ok = torch.equal(got, full)
        # equal shards, no padding argument (the bench's weak-scaling case)
        eq = gather_contacts(torch.full((3, 4), float(rank)), dist)
        ok = ok and eq.shape == (3 * world, 4) and all(torch.all(eq[3 * r:3 * r + 3] == r) for r in range(world))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_samples", [8, 7, 1])
def test_gather_contacts_world2_gloo(n_samples):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_samples, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_single_process_is_identity():
    x = torch.randn(4, 6)
    assert gather_contacts(x, None) is x
