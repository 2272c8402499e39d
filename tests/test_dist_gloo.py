"""CPU, world_size 2, gloo: the host-side sharding + index gather of the multi-GPU path (vlg/dist.py).
The decode itself is replaced by a deterministic stand-in (there is no GPU here); what is under test is that every
sample is decoded exactly once, on the right rank, with the right per-rank seed, and lands in the right row."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vlg import dist as vd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_generate(model, cond, max_new_tokens, emb_masks=None, seed=0, **kw):
    base = cond.to(torch.int32).view(-1, 1) * 1000 + seed * 100000
    return base + torch.arange(max_new_tokens, dtype=torch.int32).view(1, -1)


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cond = torch.arange(n)
        out = vd.generate_sharded(None, cond, 5, generate_fn=_fake_generate, global_seed=3)
        q.put((rank, out))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 7, 1])
def test_sharded_generate_gathers_every_sample_once(n):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = []
    for r in range(world):
        lo, hi = vd.shard_bounds(n, r, world)
        expect.append(_fake_generate(None, torch.arange(n)[lo:hi], 5, seed=3 * world + r))
    expect = torch.cat(expect)
    assert torch.equal(results[0], expect) and torch.equal(results[1], expect)
    assert results[0].dtype == torch.int32 and results[0].shape == (n, 5)


def test_shard_bounds_cover_and_partition():
    for n in (0, 1, 7, 64, 256):
        for w in (1, 2, 4, 8):
            spans = [vd.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert vd.rank_seed(5, rank=3, world_size=8) == 43      # global_seed * world + rank (train_t2i.py:70)
