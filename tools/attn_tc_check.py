#!/usr/bin/env python
"""Stand-alone parity + timing check of the tcgen05 attention core (own process under `timeout`: a trap in an
experimental kernel must not take a test session with it)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kernel_cases as kc  # noqa: E402
from oracle.cases import RPI, SYN  # noqa: E402
from sound_bubble_b200 import _abi as abi, _lib  # noqa: E402

lib = _lib.load()
for tc in (1, 0):
    abi.check(lib, lib.sb_set_option(abi.SB_OPT_ATTN_TC, tc), "opt")
    print("tcgen05 core" if tc else "SIMT core")
    print("  W=100 B=2 T=130:", kc.check_attn(lib, "cuda:0", "dis_embed", dict(SYN, use_attn=True), B=2, T=130), flush=True)
    print("  W=10  B=1 T=64 :", kc.check_attn(lib, "cuda:0", "dis_embed", dict(SYN, use_attn=True, local_atten_len=10), B=1, T=64, block=2), flush=True)
    print("  W=100 B=1 T=300:", kc.check_attn(lib, "cuda:0", "dis_embed", dict(SYN, use_attn=True), B=1, T=300, block=1), flush=True)
    print("  rpi W=50 T=200 :", kc.check_attn(lib, "cuda:0", "optim", dict(RPI, use_attn=True), B=1, T=200), flush=True)

# end to end: attention model, batch 32 x 5 s, tcgen05 core vs SIMT core (waveform RMS difference, time per call)
from bench import SYN as BSYN, radius_one_hot, synthetic_clips  # noqa: E402
from sound_bubble_b200 import Net  # noqa: E402
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = Net(**dict(BSYN, use_attn=True)).to(dev).eval()
x = synthetic_clips(32, 1234).to(dev)
inp = {"mixture": x, "dis_embed": radius_one_hot(32).to(dev)}
outs = {}
for tc in (0, 1):
    abi.check(lib, lib.sb_set_option(abi.SB_OPT_ATTN_TC, tc), "opt")
    for p_ in net._offline_pipes.values():                    # the slices' graphs were captured with the other core
        p_.close()
    net._offline_pipes.clear()
    y = net(inp)["output"]; torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); y = net(inp)["output"]; b.record(); b.synchronize()
    outs[tc] = y
    print("attention model, batch 32 x 5 s, %s core: %.1f ms" % ("tcgen05" if tc else "SIMT", a.elapsed_time(b)), flush=True)
d = outs[1] - outs[0]
print("waveform rms diff %.3g (rms %.3g), max-abs %.3g" % (float(d.pow(2).mean().sqrt()), float(outs[0].pow(2).mean().sqrt()), float(d.abs().max())))
