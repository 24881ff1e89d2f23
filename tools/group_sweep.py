#!/usr/bin/env python
"""Throughput of the pipelined streaming session when G consecutive 8 ms chunks are fed as one call (frames_per_call = G)
at several pipeline depths: the data behind the grouped throughput mode.

    python tools/group_sweep.py [G,depth ...]      e.g.  python tools/group_sweep.py 1,8 4,8 8,8 8,16
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import BATCH, CHUNK, LOOK, N_SAMPLES, SYN, T_FRAMES, radius_one_hot, synthetic_clips  # noqa: E402
from sound_bubble_b200 import Net, _lib  # noqa: E402


def main():
    combos = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(1, 8), (2, 8), (4, 8), (8, 8), (4, 16), (8, 16), (16, 8)]
    dev = torch.device("cuda", 0)
    print("CUDA_DEVICE_MAX_CONNECTIONS=%s" % os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS", "(default 8)"), flush=True)
    _lib.load()
    _lib.set_pdl(True)
    torch.manual_seed(0)
    net = Net(**SYN).to(dev).eval()
    mix = synthetic_clips(BATCH, 1234)
    dis = radius_one_hot(BATCH).to(dev)
    padded = torch.nn.functional.pad(mix, (0, LOOK)).to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    win_dev = padded.unfold(-1, CHUNK + LOOK, CHUNK).permute(2, 0, 1, 3).contiguous()          # [T, B, M, 288]
    outs = torch.empty(T_FRAMES, BATCH, 1, CHUNK, device=dev)
    ref = None
    for G, depth in combos:
        for ia, ea in ([(None, None)] if G == 1 or not os.environ.get("SWEEP_WS2") else [(None, None), (8, 7)]):
            pipe = net.streaming(BATCH, dis, pipelined=True, depth=depth, group=G, intra_algo=ia, inter_algo=ea)

            def one_pass():
                pipe.reset(); pipe.begin()
                t0 = time.perf_counter()
                for t in range(T_FRAMES):
                    pipe.feed(win_dev[t], out=outs[t])
                dt = time.perf_counter() - t0
                pipe.end()
                return dt
            for _ in range(2):
                one_pass()
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                flush.zero_(); torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); enq = one_pass(); b.record(); b.synchronize()
                ts.append(a.elapsed_time(b))
            ms = sorted(ts)[1]
            y = outs.permute(1, 2, 0, 3).reshape(BATCH, 1, T_FRAMES * CHUNK)
            if ref is None:
                ref = y.clone()
            err = float((y - ref).abs().max())
            print("G=%2d depth=%2d in_flight=%3d intra_algo=%s inter_algo=%s  %7.2f ms  %9.0f frames/s  enqueue %.1f ms  maxabs_vs_first %.2e"
                  % (G, depth, G * depth, pipe.intra_algo, pipe.inter_algo, ms, BATCH * T_FRAMES / (ms * 1e-3), enq * 1e3, err), flush=True)
            pipe.close()
            del pipe


if __name__ == "__main__":
    main()
