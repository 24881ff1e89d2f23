#!/usr/bin/env python
"""A/B of the tensor-core conv-in (SB_OPT_FRONT_TC, conv_in_tc_kernel) against conv_in_kernel on the same inputs: max difference
of the outputs and of the new history, time per launch (L2 flushed) at the grouped-streaming and offline sizes."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.cases import SYN
from sound_bubble_b200 import Net, _abi as abi, _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = Net(**SYN).to(dev).eval()
pk = net.engine().packed
cfg = net.cfg if hasattr(net, "cfg") else None
F, Cin, C = 145, 27, 32


def timeit(fn, reps=10):
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]


for (B, T) in ((2, 5), (32, 32), (32, 625)):
    g = torch.Generator().manual_seed(B * 100 + T)
    feats = torch.randn(B, T, F, Cin, generator=g).to(dev)
    cb_in = torch.randn(B, Cin, 2, F, generator=g).to(dev)
    res = []
    for opt in (0, 1):
        abi.check(lib, lib.sb_set_option(abi.SB_OPT_FRONT_TC, opt), "sb_set_option")
        x = torch.full((B, T, F, C), float("nan"), device=dev)
        cb_out = torch.full_like(cb_in, float("nan"))
        a = abi.ConvInArgs()
        a.feats, a.conv_buf_in, a.conv_buf_out = feats.data_ptr(), cb_in.data_ptr(), cb_out.data_ptr()
        a.w_pack, a.bias, a.ln_g, a.ln_b = pk.ptr("conv_w_pack"), pk.ptr("conv_bias"), pk.ptr("conv_ln_g"), pk.ptr("conv_ln_b")
        a.x = x.data_ptr()
        a.B, a.T, a.F, a.Cin, a.C = B, T, F, Cin, C
        st = torch.cuda.current_stream().cuda_stream
        fn = lambda: abi.check(lib, lib.sb_conv_in_fwd(ctypes.byref(a), st), "sb_conv_in_fwd")
        fn(); torch.cuda.synchronize()
        res.append((x.clone(), cb_out.clone(), timeit(fn)))
    (x0, c0, t0), (x1, c1, t1) = res
    print("B=%3d T=%3d  maxabs(x_tc - x_simt) = %.3e  nan=%s  history equal=%s   simt %8.1f us   tc %8.1f us"
          % (B, T, float((x0 - x1).abs().max()), bool(torch.isnan(x1).any()), bool(torch.equal(c0, c1)), t0, t1), flush=True)
