"""Evaluation plumbing: wav I/O, the torchmetrics-semantics metrics, sample / run directory contracts
of src/test_samples.py and src/utils.py.  The GPU leg runs the whole CLI path on a run directory written in the
reference's checkpoint layout and checks the metric rows against the same metrics of the reference's golden output."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import Golden
from oracle import tfgridnet_oracle as orc
from oracle.weights import make_state_dict
from sound_bubble_b200 import evaluate as ev


def test_wav_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    pcm = rng.integers(-32768, 32767, size=(6, 1000)).astype(np.int16)
    x = pcm.astype(np.float32) / 32768.0
    p = str(tmp_path / "a.wav")
    ev.write_wav(p, x, 24000)
    y = ev.read_wav(p, 24000)
    assert y.shape == (6, 1000) and np.array_equal(y, x)            # int16 / 32768, like librosa.load (src/utils.py:137-141)
    ev.write_wav(p, x[0], 24000)
    assert ev.read_wav(p).shape == (1000,)
    with pytest.raises(ValueError):
        ev.read_wav(p, 16000)


def test_metrics_follow_torchmetrics_definitions():
    g = torch.Generator().manual_seed(0)
    t = torch.randn(2, 1, 4000, generator=g)
    n = 0.1 * torch.randn(2, 1, 4000, generator=g)
    p = 0.5 * t + n
    eps = torch.finfo(torch.float32).eps
    # SNR: 10 log10(|t|^2 / |t - p|^2)
    want = 10 * torch.log10((t.pow(2).sum(-1) + eps) / ((t - p).pow(2).sum(-1) + eps))
    assert torch.allclose(ev.snr(p, t), want)
    # SI-SDR is invariant to the scale of the estimate, SNR is not
    assert torch.allclose(ev.si_sdr(p, t), ev.si_sdr(3.0 * p, t), atol=1e-4)
    a = (p * t).sum(-1, keepdim=True) / t.pow(2).sum(-1, keepdim=True)
    want = 10 * torch.log10((a * t).pow(2).sum(-1) / (a * t - p).pow(2).sum(-1))
    assert torch.allclose(ev.si_sdr(p, t), want, atol=1e-4)
    # SI-SNR = SI-SDR after mean removal
    assert torch.allclose(ev.si_snr(p + 0.3, t), ev.si_sdr(p - p.mean(-1, keepdim=True), t - t.mean(-1, keepdim=True)), atol=1e-4)
    # decay: energy ratio in dB
    assert torch.allclose(ev.compute_decay(0.1 * t, t), torch.full((2,), 20.0), atol=1e-4)
    row = ev.clip_metrics(p[0], t[0], (t + n)[0], 1)
    assert set(row) == {"n_tgt_speakers", "input_snr", "snri", "input_sisnr", "sisnri", "input_sisdr", "sisdri"}
    assert set(ev.clip_metrics(p[0], t[0], t[0], 0)) == {"n_tgt_speakers", "decay"}


def _sample_dir(root, name, mixture, voices, real=False):
    d = os.path.join(root, name)
    os.makedirs(d)
    ev.write_wav(os.path.join(d, "mixture.wav"), mixture, 24000)
    meta = {"real": real, "room": "r", "room_info": {}}
    for i, (dis, wavf) in enumerate(voices):
        meta["voice%02d" % i] = {"dis": dis, "angle": 10.0 * i}
        ev.write_wav(os.path.join(d, "mic00_voice%02d.wav" % i), wavf, 24000)
    json.dump(meta, open(os.path.join(d, "metadata.json"), "w"))
    return d


def test_load_testcase_sums_the_voices_inside_the_bubble(tmp_path):
    rng = np.random.default_rng(1)
    q = lambda a: np.round(a * 32768) / 32768
    mix = q(rng.uniform(-0.5, 0.5, (6, 960))).astype(np.float32)
    v = [q(rng.uniform(-0.2, 0.2, 960)).astype(np.float32) for _ in range(3)]
    d = _sample_dir(str(tmp_path), "00000", mix, [(0.8, v[0]), (1.4, v[1]), (2.5, v[2])])
    meta, m, gt, tgt, spatial = ev.load_testcase(d, 1.0)
    assert np.array_equal(m, mix) and np.allclose(gt[0], v[0]) and len(tgt) == 1
    _, _, gt, tgt, spatial = ev.load_testcase(d, 1.5)
    assert np.allclose(gt[0], v[0] + v[1], atol=1e-7) and spatial == {"dis_near": [0.8, 1.4], "dis_far": [2.5]}
    d2 = _sample_dir(str(tmp_path), "00001", mix, [(80, v[0]), (250, v[1])], real=True)      # real recordings: centimetres
    assert len(ev.load_testcase(d2, 1.0)[3]) == 1
    assert len(ev.load_testcase(d2, 2.0)[3]) == 1


def _run_dir(root, g):
    """config.json + checkpoints/best.pt in the reference's layout (hl_module.py:141-156, train_pt.py:96-102)."""
    rd = os.path.join(root, "run")
    os.makedirs(os.path.join(rd, "checkpoints"))
    model = "src.models.tfgridnet_realtime_clean_dis_embd3.net.Net" if g.variant == "dis_embed" else \
        "src.models.tfgridnet_realtime_clean_optim.net.Net"
    json.dump({"pl_module": "src.hl_modules.distance_based_hl_module.PLModule",
               "pl_module_args": {"model": model, "model_params": g.kwargs}}, open(os.path.join(rd, "config.json"), "w"))
    sd = make_state_dict(orc.OracleConfig.from_kwargs(g.variant, **g.kwargs), g.meta["seed"])
    torch.save({"model": sd, "optimizer": {}, "current_epoch": 3, "metric_values": {}, "statistics": {}, "scheduler": {}},
               os.path.join(rd, "checkpoints", "best.pt"))
    return rd


def test_run_dir_contract_without_a_gpu(tmp_path):
    g = Golden("wav_syn_1m")
    net, params = ev.load_run_dir(_run_dir(str(tmp_path), g), device="cpu")
    assert type(net).__module__ == "sound_bubble_b200.tfgridnet_realtime_clean_dis_embd3.net"
    assert params["pl_module_args"]["model_params"] == g.kwargs
    with pytest.raises(FileNotFoundError):
        os.remove(os.path.join(tmp_path, "run", "checkpoints", "best.pt"))
        ev.load_run_dir(os.path.join(tmp_path, "run"), device="cpu")


@pytest.mark.gpu
def test_cli_path_against_the_reference_golden_output(tmp_path):
    """The fixture clip (test_samples/syn_1m) through run dir -> Net -> batched run_testcases: the metric rows equal the
    same metrics computed on the UNMODIFIED reference's output for that clip (tests/golden/wav_syn_1m.npz)."""
    g = Golden("wav_syn_1m")
    rd = _run_dir(str(tmp_path), g)
    mix = g.mixture[0].numpy()
    rng = np.random.default_rng(2)
    voice = (np.round(0.3 * mix[0] * 32768 + rng.integers(-50, 50, mix.shape[-1])) / 32768).astype(np.float32)
    root = os.path.join(tmp_path, "samples")
    os.makedirs(root)
    dirs = [_sample_dir(root, "00000", mix, [(0.7, voice), (3.0, voice)]),           # one speaker inside the 1 m bubble
            _sample_dir(root, "00001", mix, [(1.8, voice)])]                         # nobody inside: decay only
    net, _ = ev.load_run_dir(rd, device="cuda:0")
    rows = ev.run_testcases(net, dirs, distance_threshold=1.0)
    assert [r["sample"] for r in rows] == ["00000", "00001"]
    ref_out = g.output[0]                                                           # [1, N] reference output, radius 1 m
    want0 = ev.clip_metrics(ref_out, torch.from_numpy(voice)[None], g.mixture[0, 0:1], 1)
    want1 = ev.clip_metrics(ref_out, torch.zeros(1, mix.shape[-1]), g.mixture[0, 0:1], 0)
    for got, want in ((rows[0], want0), (rows[1], want1)):
        for k, v in want.items():
            assert abs(got[k] - v) <= 0.05, (k, got[k], v)                          # the north star's 0.05 dB SI-SDR bar
    ev.main([root, rd, "--distance_threshold", "1", "--csv", os.path.join(tmp_path, "out.csv")])
    assert open(os.path.join(tmp_path, "out.csv")).readline().strip().split(",")[0] in ("decay", "input_sisdr")
