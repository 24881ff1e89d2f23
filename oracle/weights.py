"""Deterministic parity weights in the reference checkpoint layout.  TEST INFRASTRUCTURE ONLY.

The pretrained ``TFG_S_*`` checkpoint is a Google-Drive download (/root/reference/README.md:37-42) and is not
available, so parity is established on seeded weights (SURVEY.md §8c "Weights").  Every tensor is drawn from its
own ``torch.Generator`` so the set does not depend on module construction order; LayerNorm gains/biases and all
biases are made non-trivial so affine bugs cannot hide behind ones/zeros.
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict

import torch

from .tfgridnet_oracle import OracleConfig, param_shapes, stft_basis


def _fan_in(name: str, shape) -> int:
    if "rnn" in name:
        return shape[0] // 4                      # torch.nn.LSTM: U(-1/sqrt(H), 1/sqrt(H))
    if len(shape) == 1:
        return 0
    n = 1
    for d in shape[1:]:
        n *= d
    if name.endswith("deconv.weight"):            # ConvTranspose: fan_in counts the output-channel axis
        n = shape[0] * (shape[2] if len(shape) == 3 else shape[2] * shape[3])
    return n


def make_state_dict(cfg: OracleConfig, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    sd = OrderedDict()
    for idx, (name, shape) in enumerate(param_shapes(cfg).items()):
        if name.endswith("_filters"):
            sd[name] = stft_basis(cfg.n_fft, cfg.stft_chunk_size)
            continue
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        leaf = name.rsplit(".", 1)[-1]
        is_norm = ".norm." in name or "dis_norm" in name or name.endswith(("conv.1.weight", "conv.1.bias")) \
            or "dis_embedding.1" in name
        if is_norm:
            t = 0.1 * torch.randn(shape, generator=g)
            if leaf == "weight":
                t = t + 1.0
        elif name.endswith(("act.weight", ".1.weight")) and shape == (1,):      # PReLU slope
            t = torch.full(shape, 0.25) + 0.05 * torch.randn(shape, generator=g)
        elif len(shape) == 1 and "rnn" not in name:
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            k = 1.0 / max(_fan_in(name, shape), 1) ** 0.5
            t = (torch.rand(shape, generator=g) * 2.0 - 1.0) * k
        sd[name] = t.float().contiguous()
    return sd


def state_dict_digest(sd) -> str:
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


def synthetic_mixture(batch: int, n_mics: int, n_samples: int, seed: int = 1234, correlated: bool = True):
    """Synthetic clips of SURVEY.md §8d: 0.1*randn, optionally one delayed common source per clip so that the
    inter-microphone level/phase features are not degenerate."""
    g = torch.Generator().manual_seed(seed)
    if not correlated:
        return 0.1 * torch.randn(batch, n_mics, n_samples, generator=g)
    src = 0.1 * torch.randn(batch, n_samples + 8, generator=g)
    noise = 0.02 * torch.randn(batch, n_mics, n_samples, generator=g)
    gains = 0.5 + torch.rand(batch, n_mics, generator=g)
    out = torch.empty(batch, n_mics, n_samples)
    for m in range(n_mics):
        d = (3 * m + 1) % 9
        out[:, m] = gains[:, m:m + 1] * src[:, d:d + n_samples]
    return out + noise


def radius_one_hot(batch: int) -> torch.Tensor:
    """dis_embed cycling 1 m / 1.5 m / 2 m  (/root/reference/src/test_samples.py:96-102)."""
    table = torch.tensor([[0., 0., 1.], [0., 1., 0.], [1., 0., 0.]])
    return table[torch.arange(batch) % 3].clone()
