#!/usr/bin/env python
"""Data-parallel training step of the TFG_S separator, one process per GPU (BASELINE config 4's shape):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/train_ddp.py [--batch-per-gpu 4] [--seconds 2] [--steps 3]

Each rank runs forward + backward of its shard through the hand-written kernels (Net in train() mode), then ONE NCCL
all-reduce over the flat 2 MB gradient buffer (train_dist.FlatGradReducer), clip after the reduction, Adam - the order
PLModule.backprop uses (src/hl_modules/distance_based_hl_module.py:433-441).  Checks that the reduced gradient equals the
gradient of the global-batch mean loss computed on one GPU, then times the step (device time, max over ranks).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import SYN, radius_one_hot  # noqa: E402


def neg_snr(est, tgt):      # asteroid SingleSrcNegSDR('snr') as used by src/losses/SNRLosses.py:12-13, mean over the batch
    return -(10 * torch.log10(tgt.pow(2).sum(-1) / ((est - tgt).pow(2).sum(-1) + 1e-8))).mean()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch-per-gpu", type=int, default=4)
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from sound_bubble_b200 import Net
    from sound_bubble_b200.train_dist import FlatGradReducer, backprop
    torch.manual_seed(0)
    net = Net(**SYN).to(dev).train()
    n = int(args.seconds * 24000) // 192 * 192
    B = args.batch_per_gpu * world
    g = torch.Generator().manual_seed(1)
    mix = 0.1 * torch.randn(B, 6, n, generator=g)
    tgt = 0.1 * torch.randn(B, 1, n, generator=g)
    dis = radius_one_hot(B)
    lo, hi = rank * args.batch_per_gpu, (rank + 1) * args.batch_per_gpu
    my = [t[lo:hi].to(dev) for t in (mix, tgt, dis)]
    red = FlatGradReducer(net.parameters())
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)

    # (1) the reduced gradient is the global-batch gradient
    red.zero_grad()
    neg_snr(net({"mixture": my[0], "dis_embed": my[2]})["output"], my[1]).backward()
    red.all_reduce_mean()
    reduced = red.flat.clone()
    err = 0.0
    if rank == 0:
        red.zero_grad()
        neg_snr(net({"mixture": mix.to(dev), "dis_embed": dis.to(dev)})["output"], tgt.to(dev)).backward()
        err = float((red.flat - reduced).abs().max() / red.flat.abs().max())

    # (2) timed steps
    def step():
        red.zero_grad()
        loss = neg_snr(net({"mixture": my[0], "dis_embed": my[2]})["output"], my[1])
        loss.backward()
        backprop(red, opt, grad_clip=1.0)
        return loss
    step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        frames = B * (n // 192)
        print(json.dumps({"what": "data-parallel training step, TFG_S, fp32, one flat-gradient all-reduce per step", "n_gpus": world,
                          "global_batch": B, "seconds": args.seconds, "ms_per_step": float(ms), "train_frames_per_s": frames / float(ms) * 1e3,
                          "reduced_vs_global_batch_grad_relerr": err, "grad_floats": red.numel, "loss": float(loss.detach())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
