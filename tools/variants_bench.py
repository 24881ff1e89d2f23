#!/usr/bin/env python
"""Context numbers for the other shipped configurations (and the dormant attention branch): per-chunk latency of the
in-order streaming session, pipelined streaming throughput and whole-utterance throughput at batch 32 x 5 s."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import SYN, radius_one_hot, synthetic_clips, windows_of  # noqa: E402
from oracle.cases import OPI, RPI  # noqa: E402  (configuration dictionaries only)
from sound_bubble_b200 import Net, NetOptim  # noqa: E402

dev = torch.device("cuda", 0)
B, T = 32, 625
mix = synthetic_clips(B, 1234)
win = windows_of(mix).to(dev)
x = mix.to(dev)
dis = radius_one_hot(B).to(dev)
out = torch.empty(T, B, 1, 192, device=dev)


def timed(fn, n=3):
    fn(); fn(); torch.cuda.synchronize()            # two warm-up calls: lazy kernel loading, pipe creation, allocator
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n


for name, cls, kw in (("TFG_S (syn, D32 B6)", Net, SYN), ("TFG_S + attention (L4 E2 W100)", Net, dict(SYN, use_attn=True)),
                      ("Orange-Pi (optim, D32 B6)", NetOptim, OPI), ("Raspberry-Pi (optim, D16 B3 conv-LSTM k5)", NetOptim, RPI)):
    torch.manual_seed(0)
    net = cls(**kw).to(dev).eval()
    sess = net.streaming(B, dis)

    def in_order():
        sess.reset()
        for t in range(T):
            sess.feed(win[t])
    ms_io = timed(in_order)
    pipe = net.streaming(B, dis, pipelined=True)

    def pipelined():
        pipe.reset(); pipe.begin()
        for t in range(T):
            pipe.feed(win[t], out[t])
        pipe.end()
    ms_p = timed(pipelined)
    inp = {"mixture": x, "dis_embed": dis}
    ms_off = timed(lambda: net(inp))
    n_par = sum(p.numel() for p in net.parameters())
    print("%-44s %7d params | in-order %6.1f us/chunk | pipelined %7.0f frames/s | offline %7.0f frames/s (%.1f ms)"
          % (name, n_par, 1e3 * ms_io / T, B * T / ms_p * 1e3, B * T / ms_off * 1e3, ms_off), flush=True)
    pipe.close()
