"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch, run their shard through a stand-in for the
per-rank forward, and gather; the union must equal the single-process result (no data-path collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sound_bubble_b200.dist import gather_outputs, shard_bounds, shard_inputs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_net(inputs):
    """Row-independent stand-in for the separator (the real one needs a GPU): per-utterance function of its inputs."""
    x = inputs["mixture"]
    return {"output": (x.sum(1, keepdim=True) * inputs["dis_embed"].argmax(1).view(-1, 1, 1).float()).cumsum(-1)}


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        inputs = {"mixture": torch.randn(n_items, 6, 50, generator=g),
                  "dis_embed": torch.eye(3)[torch.arange(n_items) % 3], "tag": "x"}
        local = _fake_net(shard_inputs(inputs, world, rank))["output"]
        lo, hi = shard_bounds(n_items, world, rank)
        assert local.shape[0] == hi - lo
        full = gather_outputs(local, n_items)
        ref = _fake_net(inputs)["output"]
        only0 = gather_outputs(local, n_items, dst=0)
        ok = bool(torch.equal(full, ref)) and ((only0 is None) == (rank != 0))
        # the timing reduction bench.py uses: max over ranks
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and float(t) == float(world)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 5, 1])
def test_two_rank_sharding_matches_single_process(n_items):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 5, 32, 256, 257):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= (n + w - 1) // w
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)
