"""Runs the UNMODIFIED reference module out of oracle/_ref (see oracle/build_ref.py).  BASELINE INFRASTRUCTURE ONLY:
used by bench.py's `--impl reference` arm / cpu_baseline leg and by tests; never by the product package."""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
SHIMS = os.path.join(ROOT, "oracle", "shims")


def available() -> bool:
    """True when oracle/_ref holds the files its manifest lists, unmodified."""
    try:
        man = json.load(open(os.path.join(REF_DIR, "MANIFEST.json")))
        return all(hashlib.sha256(open(os.path.join(REF_DIR, rel), "rb").read()).hexdigest() == h
                   for rel, h in man["files"].items())
    except Exception:
        return False


def reference_net(kwargs: dict, state_dict=None, variant: str = "dis_embed"):
    """The reference's own Net (DE3/net.py:21 or OPT/net.py), weights loaded strictly."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` in the build container")
    for p in (SHIMS, REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    if variant == "dis_embed":
        from src.models.tfgridnet_realtime_clean_dis_embd3.net import Net
    else:
        from src.models.tfgridnet_realtime_clean_optim.net import Net
    net = Net(**kwargs).eval()
    if state_dict is not None:
        # the stand-in STFTFB of oracle/shims does not register asteroid's `torch_window` buffer (nothing reads it);
        # every other key must match exactly
        own = net.state_dict()
        sd = {k: v.detach().cpu() for k, v in state_dict.items() if not (k.endswith("filterbank.torch_window") and k not in own)}
        net.load_state_dict(sd, strict=True)
    return net


def streaming_sample(net, mixture: torch.Tensor, dis_embed: torch.Tensor, chunk: int, pad: int, n_chunks: int, warm: int = 1):
    """edge/causal_infer.py:28-47 at the batch of `mixture` [B, M, >= chunk*(n_chunks+warm)+pad]: the window rolls left by
    one chunk per call, pad=False, state threaded through.  Returns (seconds for the timed n_chunks, outputs [B,S,n*chunk])."""
    B, M, _ = mixture.shape
    state = net.init_buffers(B, "cpu")
    frame = torch.zeros(B, M, chunk + pad)
    frame[..., -pad:] = mixture[..., :pad]
    outs, t0 = [], None
    with torch.no_grad():
        for k in range(n_chunks + warm):
            if k == warm:
                t0 = time.perf_counter()
            i = pad + k * chunk
            frame = torch.roll(frame, shifts=-chunk, dims=-1)
            frame[..., -chunk:] = mixture[..., i:i + chunk]
            r = net({"mixture": frame, "dis_embed": dis_embed}, state, pad=False)
            state = r["next_state"]
            outs.append(r["output"])
    return time.perf_counter() - t0, torch.cat(outs, dim=-1)
