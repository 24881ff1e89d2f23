"""Device batch assembly: kernel vs the CPU restatement of the reference's dataset + perturbations
(host-emulated build here, the real library on the GPU), and the restatement vs the reference's own perturbation classes."""
import importlib.util
import os
import sys

import pytest
import torch

from oracle import batch_oracle as bo
from sound_bubble_b200.batching import prepare_batch

REF = "/root/reference"


def _case(B, M, V, N, seed, full=True):
    g = torch.Generator().manual_seed(seed)
    mix = torch.randint(-20000, 20000, (B, M, N), generator=g, dtype=torch.int16)
    voices = torch.randint(-8000, 8000, (B, V, N), generator=g, dtype=torch.int16) if V else None
    inside = (torch.rand(B, V, generator=g) < 0.6).to(torch.uint8) if V else None
    kw = {"radius_idx": torch.randint(0, 3, (B,), generator=g, dtype=torch.int32)}
    if full:
        kw["gain"] = 10 ** ((2 * (torch.rand(B, M, generator=g) - 0.5) * 6.0) / 20)
        kw["shift"] = torch.randint(-5, 6, (B, M), generator=g, dtype=torch.int32)
        kw["shift"][0, 0] = -(N + 3)                                           # more than one period
        drop = torch.zeros(B, M, dtype=torch.uint8)
        drop[:, 1:] = (torch.rand(B, M - 1, generator=g) < 0.3).to(torch.uint8)
        kw["drop"] = drop
        kw["peak_scale"] = torch.where(torch.rand(B, generator=g) < 0.7, 0.2 + torch.rand(B, generator=g), torch.zeros(B))
    return mix, voices, inside, kw


def _check(lib, dev, B, M, V, N, seed, full=True):
    mix, voices, inside, kw = _case(B, M, V, N, seed, full)
    ref_mix, ref_tgt, ref_dis = bo.assemble(mix, voices, inside, **kw)
    to = lambda t: None if t is None else t.to(dev)
    inputs, tgt = prepare_batch(to(mix), to(voices), to(inside), lib=lib, **{k: to(v) for k, v in kw.items()})
    if dev != "cpu":
        torch.cuda.synchronize()
    assert torch.equal(inputs["dis_embed"].cpu(), ref_dis)
    assert float((inputs["mixture"].cpu() - ref_mix).abs().max()) <= 2e-6 * max(1.0, float(ref_mix.abs().max()))
    assert float((tgt.cpu() - ref_tgt).abs().max()) <= 2e-6 * max(1.0, float(ref_tgt.abs().max()))


@pytest.mark.parametrize("B,M,V,N,full", [(2, 6, 3, 4096, True), (3, 6, 2, 5003, True), (1, 2, 0, 2048, False),
                                          (2, 6, 1, 777, True)])
def test_kernel_logic_on_the_host_emulated_build(B, M, V, N, full):
    from emu.emu_lib import load
    _check(load(), "cpu", B, M, V, N, seed=B * 100 + V, full=full)


@pytest.mark.gpu
@pytest.mark.parametrize("B,M,V,N,full", [(4, 6, 3, 120000, True), (3, 6, 2, 120001, True), (2, 6, 0, 48000, False),
                                          (32, 6, 4, 120000, True)])
def test_kernel_on_the_gpu(B, M, V, N, full):
    from sound_bubble_b200 import _lib
    _check(_lib.load(), "cuda:0", B, M, V, N, seed=7 + B, full=full)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only mounted in the build container")
def test_restatement_matches_the_reference_perturbation_classes():
    def load(name):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, "src/datasets/perturbations", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return getattr(mod, name)
    Shift, Gain, Drop, Peak = (load(n) for n in ("SampleShiftPerturbation", "ChannelGainPerturbation",
                                                   "ChannelDropPerturbation", "PeakNormPerturbation"))
    M, N = 6, 3000
    mix, voices, inside, _ = _case(1, M, 2, N, seed=3, full=False)
    inside[:] = 1
    audio0, gt0 = bo.pcm(mix[0]).clone(), (bo.pcm(voices[0, 0]) + bo.pcm(voices[0, 1]))[None].clone()
    # run the reference classes with a known seed, then replay their draws (same order of torch RNG calls)
    torch.manual_seed(5)
    a, t = Shift(4)(audio0.clone(), gt0.clone())
    a, t = Gain(6.0)(a, t)
    a, t = Drop(2)(a, t)
    a, t = Peak(0.3, 0.9)(a, t)
    torch.manual_seed(5)
    shift = torch.tensor([[int(torch.randint(-4, 5, (1,))) for _ in range(M)]], dtype=torch.int32)
    gain = torch.tensor([[10 ** ((2 * (torch.rand((1,)).item() - 0.5) * 6.0) / 20) for _ in range(M)]])
    n_drop = torch.randint(1, 3, (1,)).item()
    perm = 1 + torch.randperm(M - 1)
    drop = torch.zeros(1, M, dtype=torch.uint8)
    drop[0, perm[:n_drop]] = 1
    scale = torch.tensor([torch.randn((1,)).item() * (0.9 - 0.3) + 0.3])
    m2, t2, _ = bo.assemble(mix, voices, inside, None, gain, shift, drop, scale)
    assert float((m2[0] - a).abs().max()) <= 1e-6 and float((t2[0] - t).abs().max()) <= 1e-6
