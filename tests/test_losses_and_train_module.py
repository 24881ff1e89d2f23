"""The loss restatements (sound_bubble_b200/losses.py) and the PLModule-compatible per-rank harness
(train_dist.TrainModule) on CPU: closed forms, the reference wrappers' contracts, and - gloo, world size 2 - three
optimizer steps on two shards reproduce the single-process loss curve of the global batch."""
import json
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from sound_bubble_b200 import losses as L
from sound_bubble_b200.train_dist import TrainModule, import_attr

HERE = os.path.dirname(os.path.abspath(__file__))


def _pair(n=3, t=4000, seed=0):
    g = torch.Generator().manual_seed(seed)
    gt = torch.randn(n, 1, t, generator=g)
    return 0.7 * gt + 0.3 * torch.randn(n, 1, t, generator=g), gt          # an estimate a few dB above the noise


def test_single_src_neg_sdr_closed_forms():
    est, gt = _pair()
    e, g = est[:, 0] - est[:, 0].mean(-1, keepdim=True), gt[:, 0] - gt[:, 0].mean(-1, keepdim=True)
    snr = 10 * torch.log10(g.pow(2).sum(-1) / (e - g).pow(2).sum(-1))
    assert torch.allclose(L.SingleSrcNegSDR("snr")(est[:, 0], gt[:, 0]), -snr, atol=1e-4)
    a = (e * g).sum(-1, keepdim=True) / g.pow(2).sum(-1, keepdim=True)
    sisdr = 10 * torch.log10((a * g).pow(2).sum(-1) / (e - a * g).pow(2).sum(-1))
    assert torch.allclose(L.SingleSrcNegSDR("sisdr")(est[:, 0], gt[:, 0]), -sisdr, atol=1e-4)
    sdsdr = 10 * torch.log10((a * g).pow(2).sum(-1) / (e - g).pow(2).sum(-1))
    assert torch.allclose(L.SingleSrcNegSDR("sdsdr")(est[:, 0], gt[:, 0]), -sdsdr, atol=1e-4)
    # scale invariance of SI-SDR, and none for SNR
    assert torch.allclose(L.SingleSrcNegSDR("sisdr")(3 * est[:, 0], gt[:, 0]), -sisdr, atol=1e-4)
    with pytest.raises(TypeError):
        L.SingleSrcNegSDR("snr")(est, gt)                       # [batch, time] only, as asteroid


def test_reference_wrappers():
    est, gt = _pair()
    gt[1] = 0                                                   # nobody inside the bubble: L1 on the output x neg_weight
    est.requires_grad_(True)
    v = L.SNRLPLoss("snr", neg_weight=100)(est, gt)
    assert v.shape == (3,)
    assert torch.allclose(v[1], 100 * est[1].abs().mean(), rtol=1e-5)
    assert torch.allclose(v[[0, 2]], L.SingleSrcNegSDR("snr")(est[[0, 2], 0], gt[[0, 2], 0]), atol=1e-5)
    v.mean().backward()
    assert bool(torch.isfinite(est.grad).all()) and float(est.grad[1].abs().max()) > 0
    for name in ("sisdr", "snr", "fused", "max_fused", "sdsdr", "full"):
        assert L.SNRLosses(name)(est.detach(), gt).shape == (3,)
    with pytest.raises(AssertionError):
        L.SNRLosses("nope")


def test_multi_resolution_stft_loss():
    est, gt = _pair(2, 24000)
    kw = dict(l1_ratio=10, sample_rate=24000, perceptual_weighting=True, w_sc=0, w_log_mag=0, w_lin_mag=20)     # finetune JSONs
    m = L.MultiResoFuseLoss(**kw)
    assert float(m(gt, gt)) < 1e-4                              # identical signals
    est.requires_grad_(True)
    v = m(est, gt)
    v.backward()
    assert v.ndim == 0 and bool(torch.isfinite(v)) and bool(torch.isfinite(est.grad).all())
    # the loss is the mean over three resolutions of 20 x L1(|STFT|) on A-weighted signals + 10 x L1
    k = L.a_weighting_fir(24000)
    assert k.shape == (1, 1, 101)
    fa = torch.nn.functional.conv1d(est.detach().reshape(2, 1, -1), k, padding=50)
    fb = torch.nn.functional.conv1d(gt.reshape(2, 1, -1), k, padding=50)
    acc = 0.0
    for n, h, w in ((1024, 120, 600), (2048, 240, 1200), (512, 50, 240)):
        sa = torch.stft(fa[:, 0], n, h, w, torch.hann_window(w), return_complex=True).abs()
        sb = torch.stft(fb[:, 0], n, h, w, torch.hann_window(w), return_complex=True).abs()
        acc = acc + 20 * (sa - sb).abs().mean()
    want = acc / 3 + 10 * (est.detach() - gt).abs().mean()
    assert abs(float(v) - float(want)) <= 1e-3 * float(want)
    # A-weighting: ~0 dB at 1 kHz, strong attenuation at 50 Hz
    t = torch.arange(24000) / 24000.0
    gain = lambda f: float(torch.nn.functional.conv1d(torch.sin(2 * torch.pi * f * t).view(1, 1, -1), k, padding=50)[0, 0, 2000:-2000].pow(2).mean().sqrt() * 2 ** 0.5)  # noqa: E731
    assert abs(gain(1000.0) - 1.0) < 0.1 and gain(50.0) < 0.2
    # default auraloss terms (spectral convergence + log magnitude) are implemented as well
    assert bool(torch.isfinite(L.MultiResolutionSTFTLoss()(est.detach(), gt)))


def test_experiment_fixtures_resolve():
    """Every shipped experiment JSON's pl_module_args construct the harness (model -> this package, loss -> restatement)."""
    exps = json.load(open(os.path.join(HERE, "golden", "experiments.json")))
    assert len(exps) == 6
    for name, e in exps.items():
        a = e["pl_module_args"]
        assert import_attr(a["model"]).__module__.startswith("sound_bubble_b200.")
        loss = import_attr(a["loss"])(**a["loss_params"])
        est, gt = _pair(2, 24000)
        assert bool(torch.isfinite(loss(est=est, gt=gt).mean())), name


# ---- gloo: TrainModule on two shards == one process on the global batch -------------------------------------------------
class TinySeparator(nn.Module):
    """Same call contract as Net: dict in, {'output': [B, 1, N]} out."""

    def __init__(self, ch=3):
        super().__init__()
        self.conv = nn.Conv1d(ch, 8, 5, padding=2)
        self.rnn = nn.LSTM(8, 8, batch_first=True)
        self.out = nn.Conv1d(8, 1, 1)

    def forward(self, inputs, input_state=None, pad=True):
        y = torch.tanh(self.conv(inputs["mixture"]))
        y = self.rnn(y.transpose(1, 2))[0].transpose(1, 2)
        return {"output": self.out(y), "next_state": None}


ARGS = dict(model=TinySeparator, model_params={}, sr=24000, optimizer="torch.optim.Adam", optimizer_params={"lr": 1.2e-3},
            scheduler="torch.optim.lr_scheduler.ReduceLROnPlateau", scheduler_params={"mode": "min", "patience": 8, "factor": 0.5, "min_lr": 1e-6},
            loss="src.losses.SNRLP.SNRLPLoss", loss_params={"snr_loss_name": "snr", "neg_weight": 100},
            metrics=["snr_i", "si_snr_i", "si_sdr_i"], grad_clip=1)


def _batch(n=6, t=600):
    g = torch.Generator().manual_seed(5)
    mix = torch.randn(n, 3, t, generator=g)
    tgt = 0.5 * mix[:, :1] + 0.1 * torch.randn(n, 1, t, generator=g)
    tgt[2] = 0                                                  # one negative sample: per-item losses differ in kind
    return mix, tgt, torch.tensor([1, 1, 0, 1, 2, 1])


def _run(module, mix, tgt, nspk, steps=3):
    module.train()
    curve = []
    for i in range(steps):
        module.reset_grad()
        loss, n = module.training_step(({"mixture": mix}, {"target": tgt, "num_target_speakers": nspk}), i)
        loss.backward()
        module.backprop()
        curve.append((float(loss.detach()), n))
    return curve


def _tm_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        m = TrainModule(**ARGS)
        mix, tgt, nspk = _batch()
        lo, hi = (0, 4) if rank == 0 else (4, 6)                # ragged shards: 4 + 2 items
        curve = _run(m, mix[lo:hi], tgt[lo:hi], nspk[lo:hi])
        q.put((rank, curve, torch.cat([p.detach().reshape(-1) for p in m.model.parameters()]).tolist()))
    finally:
        dist.destroy_process_group()


def test_train_module_two_ranks_reproduce_the_global_batch_curve(tmp_path):
    torch.manual_seed(0)
    ref = TrainModule(**ARGS)
    mix, tgt, nspk = _batch()
    ref_curve = _run(ref, mix, tgt, nspk)
    ref_params = torch.cat([p.detach().reshape(-1) for p in ref.model.parameters()])
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_tm_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    for step in range(3):                                       # item-weighted mean of the shard losses == global-batch loss
        (l0, n0), (l1, n1) = res[0][1][step], res[1][1][step]
        assert abs((l0 * n0 + l1 * n1) / (n0 + n1) - ref_curve[step][0]) <= 1e-4 * max(1.0, abs(ref_curve[step][0])), step
    for _, _, params in res:                                    # both ranks hold the single-process parameters after 3 steps
        assert float((torch.tensor(params) - ref_params).abs().max()) <= 2e-5
    # PLModule's bookkeeping and checkpoint layout
    assert ref.get_avg_metric_at_epoch("train/loss") == pytest.approx(sum(c for c, _ in ref_curve) / 3, rel=1e-5)
    assert "train/si_sdr_i" in ref.metric_values[0]
    ref.log_metric("val/loss", 1.0, batch_size=2)
    path = str(tmp_path / "last.pt")
    ref.on_epoch_end(best_path=str(tmp_path / "best.pt"))
    ref.dump_state(path)
    state = torch.load(path, weights_only=False)
    assert set(state) == {"model", "optimizer", "current_epoch", "metric_values", "statistics", "scheduler"} and state["current_epoch"] == 1
    assert os.path.exists(str(tmp_path / "best.pt"))
    torch.manual_seed(1)
    other = TrainModule(**ARGS)
    other.load_state(path)
    assert other.epoch == 1 and all(torch.equal(a, b) for a, b in zip(other.model.state_dict().values(), ref.model.state_dict().values()))
