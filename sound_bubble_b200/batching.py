"""Batch assembly on the device: the sample-sized part of the reference's Dataset.__getitem__ + collate
(src/datasets/general_multisrc_dataset_dis_embed.py:112-218) and of its per-channel perturbations
(src/datasets/perturbations/*.py) as one HBM-bound kernel over PCM16 that already sits in device memory.
The random draws (which voices, shifts, gains, drops, peak scales) are the caller's: tiny host-side decisions."""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _abi as abi
from . import _lib

RADII = (1.0, 1.5, 2.0)                                    # radius_idx 0, 1, 2 -> one-hot [0,0,1], [0,1,0], [1,0,0]


def _ptr(t: Optional[torch.Tensor], dtype, shape, name, dev):
    if t is None:
        return None, None
    if t.dtype != dtype or tuple(t.shape) != tuple(shape) or t.device != dev:
        raise ValueError("%s must be %s of shape %s on %s (got %s %s on %s)" % (name, dtype, tuple(shape), dev, t.dtype,
                                                                              tuple(t.shape), t.device))
    t = t.contiguous()
    return t.data_ptr(), t


def prepare_batch(mix: torch.Tensor, voices: Optional[torch.Tensor], inside: Optional[torch.Tensor],
                  radius_idx: Optional[torch.Tensor] = None, gain: Optional[torch.Tensor] = None,
                  shift: Optional[torch.Tensor] = None, drop: Optional[torch.Tensor] = None,
                  peak_scale: Optional[torch.Tensor] = None, lib=None) -> Tuple[dict, torch.Tensor]:
    """mix int16 [B, M, N], voices int16 [B, V, N] (solo tracks at microphone 0), inside uint8 [B, V]; optional
    radius_idx int32 [B], gain float32 [B, M], shift int32 [B, M], drop uint8 [B, M], peak_scale float32 [B]
    -> ({'mixture': f32 [B, M, N], 'dis_embed': f32 [B, 3]}, target f32 [B, 1, N])."""
    lib = lib if lib is not None else _lib.load()
    if lib is _lib._cdll:
        _lib.require_cuda(mix)
    dev = mix.device
    if mix.dtype != torch.int16 or mix.dim() != 3:
        raise ValueError("mix must be int16 [B, M, N]")
    B, M, N = mix.shape
    V = 0 if voices is None else voices.shape[1]
    a = abi.PrepareArgs()
    keep = []
    a.mix, t = _ptr(mix, torch.int16, (B, M, N), "mix", dev); keep.append(t)
    if V:
        a.voices, t = _ptr(voices, torch.int16, (B, V, N), "voices", dev); keep.append(t)
        a.inside, t = _ptr(inside, torch.uint8, (B, V), "inside", dev); keep.append(t)
    a.gain, t = _ptr(gain, torch.float32, (B, M), "gain", dev); keep.append(t)
    a.shift, t = _ptr(shift, torch.int32, (B, M), "shift", dev); keep.append(t)
    a.drop, t = _ptr(drop, torch.uint8, (B, M), "drop", dev); keep.append(t)
    a.peak_scale, t = _ptr(peak_scale, torch.float32, (B,), "peak_scale", dev); keep.append(t)
    a.radius_idx, t = _ptr(radius_idx, torch.int32, (B,), "radius_idx", dev); keep.append(t)
    mixture = torch.empty(B, M, N, dtype=torch.float32, device=dev)
    target = torch.empty(B, 1, N, dtype=torch.float32, device=dev)
    dis = torch.empty(B, 3, dtype=torch.float32, device=dev)
    a.mixture, a.target, a.dis_embed = mixture.data_ptr(), target.data_ptr(), dis.data_ptr()
    if peak_scale is not None:
        ws = torch.empty(max(int(lib.sb_prepare_workspace_floats(B, M, N)), 1), dtype=torch.float32, device=dev)
        a.peak_ws = ws.data_ptr()
        keep.append(ws)
    a.B, a.M, a.V, a.N = B, M, V, N
    stream = torch.cuda.current_stream(dev).cuda_stream if mix.is_cuda else 0
    abi.check(lib, lib.sb_prepare_batch_fwd(ctypes.byref(a), stream), "sb_prepare_batch_fwd")
    return {"mixture": mixture, "dis_embed": dis}, target
