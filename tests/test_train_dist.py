"""Data-parallel training plumbing on CPU (gloo, world size 2): one all_reduce per step on a flat gradient buffer gives
the single-process global-batch gradients; clipping happens after the reduction; checkpoints keep the reference layout."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from sound_bubble_b200.dist import shard_bounds
from sound_bubble_b200.train_dist import FlatGradReducer, backprop, dump_state, load_state


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class Tiny(nn.Module):                      # the shapes of one GridNet inter path: LN -> LSTM -> Linear -> residual
    def __init__(self):
        super().__init__()
        self.norm = nn.LayerNorm(8)
        self.rnn = nn.LSTM(8, 16, batch_first=True)
        self.lin = nn.Linear(16, 8)

    def forward(self, x):
        return x + self.lin(self.rnn(self.norm(x))[0])


def _data(n):
    g = torch.Generator().manual_seed(11)
    return torch.randn(n, 20, 8, generator=g), torch.randn(n, 20, 8, generator=g)


def _single_process_reference(n, clip):
    torch.manual_seed(0)
    m = Tiny()
    x, y = _data(n)
    loss = ((m(x) - y) ** 2).mean(dim=(1, 2)).mean()            # mean over the GLOBAL batch (hl_module.py:321)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(m.parameters(), clip)
    return torch.cat([p.grad.reshape(-1) for p in m.parameters()])


def _worker(rank, world, port, n, clip, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        m = Tiny()
        red = FlatGradReducer(m.parameters())
        opt = torch.optim.SGD(m.parameters(), lr=0.0)
        calls = []
        real = dist.all_reduce
        dist.all_reduce = lambda *a, **k: (calls.append(1), real(*a, **k))[1]
        x, y = _data(n)
        lo, hi = shard_bounds(n, world, rank)
        red.zero_grad()
        loss = ((m(x[lo:hi]) - y[lo:hi]) ** 2).mean(dim=(1, 2)).mean()
        loss.backward()
        backprop(red, opt, grad_clip=clip)
        dist.all_reduce = real
        q.put((rank, red.flat.tolist(), len(calls)))
    finally:
        dist.destroy_process_group()


def test_one_all_reduce_gives_the_global_batch_gradients():
    world, n, clip = 2, 8, 0.05                                 # equal shards: mean of per-rank means == global mean
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, clip, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    ref = _single_process_reference(n, clip)
    for rank, flat, n_calls in res:
        flat = torch.tensor(flat)
        assert n_calls == 1, "exactly one collective per step"
        assert torch.allclose(flat, ref, atol=1e-6, rtol=1e-5), float((flat - ref).abs().max())
    assert float(torch.linalg.vector_norm(ref)) <= clip * 1.0001        # the clip was active in this case


def _ragged_worker(rank, world, port, n, clip, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        m = Tiny()
        red = FlatGradReducer(m.parameters())
        opt = torch.optim.SGD(m.parameters(), lr=0.0)
        x, y = _data(n)
        lo, hi = shard_bounds(n, world, rank)
        red.zero_grad()
        if hi > lo:
            ((m(x[lo:hi]) - y[lo:hi]) ** 2).mean(dim=(1, 2)).mean().backward()
        else:
            red.flat.fill_(float("nan"))                        # what a mean over an empty shard would leave behind
        backprop(red, opt, grad_clip=clip, n_local=hi - lo)
        q.put((rank, hi - lo, red.flat.tolist()))
    finally:
        dist.destroy_process_group()


def test_ragged_and_empty_shards_still_give_the_global_batch_mean():
    for n in (3, 1):                                            # 2 + 1 items, and 1 + 0 (an empty shard)
        world, clip = 2, 10.0
        port = _free_port()
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_ragged_worker, args=(r, world, port, n, clip, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = [q.get(timeout=180) for _ in range(world)]
        for p in procs:
            p.join(timeout=60)
        assert sorted(r[1] for r in res) == sorted([n - n // 2, n // 2]) or sum(r[1] for r in res) == n
        ref = _single_process_reference(n, clip)
        for rank, _, flat in res:
            flat = torch.tensor(flat)
            assert torch.isfinite(flat).all()
            assert torch.allclose(flat, ref, atol=1e-6, rtol=1e-5), (n, float((flat - ref).abs().max()))


def test_views_survive_steps_and_checkpoint_layout(tmp_path):
    torch.manual_seed(1)
    m = Tiny()
    red = FlatGradReducer(m.parameters())
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    x, y = _data(4)
    for _ in range(2):
        red.zero_grad()
        ((m(x) - y) ** 2).mean().backward()
        assert float(red.flat.abs().sum()) > 0
        backprop(red, opt, grad_clip=1.0)                       # no process group: the reduction is the identity
    red.check_views()
    path = str(tmp_path / "last.pt")
    dump_state(path, nn.DataParallel(m) if False else m, opt, epoch=3, metric_values={"val/loss": {"epoch": 1.0}})
    state = torch.load(path, weights_only=False)
    assert set(state) == {"model", "optimizer", "current_epoch", "metric_values", "statistics"}
    assert all(not k.startswith("module.") for k in state["model"])
    m2 = Tiny()
    opt2 = torch.optim.Adam(m2.parameters(), lr=1e-3)
    rest = load_state(path, m2, opt2)
    assert rest["current_epoch"] == 3 and rest["metric_values"] == {"val/loss": {"epoch": 1.0}}
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


# ---- the separator's own backward kernels under data parallelism (host-emulated TEST build of the .cu sources) ----------
def _sep_case():
    import torch.nn.functional as F  # noqa: F401
    from oracle import tfgridnet_oracle as orc
    from oracle.cases import SYN
    from oracle.weights import make_state_dict, radius_one_hot, synthetic_mixture
    from sound_bubble_b200.packing import ModelConfig
    kw = dict(SYN, B=1)
    ocfg = orc.OracleConfig.from_kwargs("dis_embed", **kw)
    sd = make_state_dict(ocfg, 0)
    n = 4
    mix = synthetic_mixture(n, 6, 192 * 2 + 96, seed=21)
    tgt = 0.1 * torch.randn(n, 1, 192 * 2, generator=torch.Generator().manual_seed(22))
    return orc, ocfg, ModelConfig(variant="dis_embed", **kw), sd, mix, radius_one_hot(n), tgt


def _sep_loss(est, tgt):                    # mean over the batch of a per-utterance loss (hl_module.py:321)
    return ((est - tgt) ** 2).mean(dim=(1, 2)).mean()


def _sep_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from emu.emu_lib import load
        from sound_bubble_b200.training import differentiable_forward
        lib = load(build=False)
        orc, ocfg, cfg, sd, mix, dis, tgt = _sep_case()
        named = {k: v.clone().requires_grad_("_filters" not in k) for k, v in sd.items()}
        params = [v for v in named.values() if v.requires_grad]
        red = FlatGradReducer(params)
        lo, hi = shard_bounds(mix.shape[0], world, rank)
        red.zero_grad()
        est = differentiable_forward(lib, cfg, named, mix[lo:hi], dis[lo:hi])
        _sep_loss(est, tgt[lo:hi]).backward()
        red.all_reduce_mean()
        q.put((rank, red.flat.tolist()))
    finally:
        dist.destroy_process_group()


def test_separator_backward_kernels_under_data_parallelism():
    """2 gloo ranks, each back-propagating its shard through the training kernels (host-emulated build), one all-reduce:
    the result is the gradient of the global-batch mean loss as autograd through the oracle computes it."""
    from emu.emu_lib import load
    load()                                                      # build once, before the ranks start
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sep_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    orc, ocfg, cfg, sd, mix, dis, tgt = _sep_case()
    leaf = {k: v.clone().requires_grad_("_filters" not in k) for k, v in sd.items()}
    out, _ = orc.core_forward(leaf, ocfg, mix, dis, orc.init_state(ocfg, mix.shape[0]))
    _sep_loss(out, tgt).backward()
    ref = torch.cat([(v.grad if v.grad is not None else torch.zeros_like(v)).reshape(-1)       # unused parameters: zeros in the flat buffer
                     for v in leaf.values() if v.requires_grad])
    for rank, flat in res:
        flat = torch.tensor(flat)
        assert flat.shape == ref.shape
        assert float((flat - ref).abs().max() / ref.abs().max()) <= 2e-5, rank
