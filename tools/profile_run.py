#!/usr/bin/env python
"""Short, profiler-friendly runs of the hot path (for ncu): a few streaming chunks and/or one offline pass.

    ncu ... python tools/profile_run.py --mode streaming --chunks 20
    ncu ... python tools/profile_run.py --mode offline --batch 32 --frames 625
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import SYN, radius_one_hot  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="streaming", choices=["streaming", "offline", "both"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--chunks", type=int, default=20)
    ap.add_argument("--frames", type=int, default=625)
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--pdl", type=int, default=0)
    ap.add_argument("--intra-algo", type=int, default=0)
    ap.add_argument("--inter-algo", type=int, default=0)
    args = ap.parse_args()
    from sound_bubble_b200 import Net, _lib
    if args.pdl:
        _lib.set_pdl(True)
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net = Net(**SYN).to(dev).eval()
    eng = net.engine()
    eng.intra_algo, eng.inter_algo = args.intra_algo, args.inter_algo
    g = torch.Generator().manual_seed(1)
    dis = radius_one_hot(args.batch).to(dev)
    if args.mode in ("streaming", "both"):
        x = (0.1 * torch.randn(args.batch, 6, 192 * args.chunks + 96, generator=g)).to(dev)
        sess = net.streaming(args.batch, dis, use_graph=bool(args.graph))
        for t in range(args.chunks):
            sess.feed(x[..., t * 192: t * 192 + 288])
        torch.cuda.synchronize()
    if args.mode in ("offline", "both"):
        x = (0.1 * torch.randn(args.batch, 6, 192 * args.frames + 96, generator=g)).to(dev)
        for _ in range(2):
            net({"mixture": x, "dis_embed": dis}, pad=False)
        torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
