#!/usr/bin/env python
"""Training step of the TFG_S separator on one B200: forward + backward through the hand-written kernels, timed with CUDA
events (BASELINE config 4's per-GPU work: the reference trains with global batch 8 on 5 s clips,
syn_experiments/pretrain_stage.json).  The CPU figure beside it is bench.py's cpu_baseline.train.

    python tools/train_bench.py [--batch 8] [--seconds 5] [--steps 3] [--cpu 1]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import SYN, radius_one_hot  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--cpu", type=int, default=0, help="ignored (the CPU figure beside the training step is bench.py's cpu_baseline.train)")
    ap.add_argument("--config", default="syn", choices=["syn", "rpi"], help="syn = TFG_S (config 4), rpi = Raspberry-Pi conv-LSTM model (config 5)")
    ap.add_argument("--ffma2", type=int, default=-1, help="0 / 1 = SB_OPT_TRAIN_FFMA2 (packed FMAs in the training GEMM kernels); -1 = library default")
    ap.add_argument("--one-row", type=int, default=0, help="1 = the first LSTM training kernels (SB_OPT_TRAIN_ONE_ROW)")
    ap.add_argument("--train-tc", type=int, default=-1, help="0 / 1 = SB_OPT_TRAIN_TC (LSTM weight gradients on tcgen05); -1 = library default")
    args = ap.parse_args()
    from sound_bubble_b200 import Net, _lib
    dev = torch.device("cuda", 0)
    if args.one_row:
        _lib.load().sb_set_option(3, 1)
    if args.ffma2 >= 0:
        _lib.load().sb_set_option(4, args.ffma2)
    if args.train_tc >= 0:
        _lib.load().sb_set_option(7, args.train_tc)
    torch.manual_seed(0)
    if args.config == "rpi":
        from oracle.cases import RPI          # configuration dictionary only
        from sound_bubble_b200.tfgridnet_realtime_clean_optim.net import Net as NetOPT
        net = NetOPT(**RPI).to(dev).train()
    else:
        net = Net(**SYN).to(dev).train()
    n = int(args.seconds * 24000) // 192 * 192
    g = torch.Generator().manual_seed(1)
    mix = (0.1 * torch.randn(args.batch, 6, n, generator=g)).to(dev)
    tgt = (0.1 * torch.randn(args.batch, 1, n, generator=g)).to(dev)
    dis = radius_one_hot(args.batch).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)

    def step():
        opt.zero_grad(set_to_none=True)
        est = net({"mixture": mix, "dis_embed": dis})["output"]
        loss = -(10 * torch.log10(tgt.pow(2).sum(-1) / ((est - tgt).pow(2).sum(-1) + 1e-8))).mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    each = []
    for _ in range(args.steps):                                 # every step timed on its own, the median reported: a step that
        e0.record()                                             # sends the caching allocator back to cudaMalloc costs 1.5x
        loss = step()
        e1.record()
        torch.cuda.synchronize()
        each.append(e0.elapsed_time(e1))
    ms = sorted(each)[len(each) // 2]
    frames = args.batch * (n // 192)
    res = {"what": "training step (forward + backward + clip + Adam), %s, fp32" % ("TFG_S" if args.config == "syn" else "Raspberry-Pi conv-LSTM model"), "batch": args.batch, "seconds": args.seconds, "one_row_kernels": bool(args.one_row), "ffma2_gemms": args.ffma2, "train_tc": args.train_tc,
           "ms_per_step": ms, "ms_each_step": each, "train_frames_per_s": frames / ms * 1e3, "clips_per_s": args.batch / ms * 1e3,
           "launches_per_step": (_lib.launch_count() - l0) / args.steps,
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "loss": float(loss)}
    # forward-only and backward-only split
    torch.cuda.synchronize()
    e0.record()
    est = net({"mixture": mix, "dis_embed": dis})["output"]
    e1.record()
    l = est.pow(2).mean()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    l.backward()
    e3.record()
    torch.cuda.synchronize()
    res["fwd_ms"], res["bwd_ms"] = e0.elapsed_time(e1), e2.elapsed_time(e3)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
