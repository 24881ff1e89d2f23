#!/usr/bin/env python
"""Stand-alone parity + timing check of the warp-specialised tcgen05 + TMA LSTM kernel (SB_ALGO_TCP), run in its own
process under `timeout` so that a trap in the kernel cannot take a test session with it.
   parity : stage checks against the CPU oracle (same cases as SB_ALGO_TC's), golden fixtures end to end
   bits   : SB_ALGO_TCP against SB_ALGO_TC on identical inputs (same operand images and summation order: expected equal)
   time   : both kernels at the streaming-group and offline shapes"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kernel_cases as kc  # noqa: E402
from oracle.cases import SYN  # noqa: E402
from sound_bubble_b200 import Net, _abi as abi, _lib  # noqa: E402

lib = _lib.load()
TC, TCP = abi.SB_ALGO_TC, abi.SB_ALGO_TCP
which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda", 0)


def intra_call(pk, x, film, algo, block=1):
    B, T, F, C = x.shape
    yf, yb = torch.full_like(x, float("nan")), torch.full_like(x, float("nan"))
    a = abi.IntraArgs()
    a.x, a.y_fwd, a.y_bwd = x.data_ptr(), yf.data_ptr(), yb.data_ptr()
    if film is not None:
        a.film_scale, a.film_shift = film[0].data_ptr(), film[1].data_ptr()
    a.dir[0], a.dir[1] = pk.lstm_dir(block, "intra0"), pk.lstm_dir(block, "intra1")
    a.B, a.T, a.F, a.C, a.H, a.algo = B, T, F, C, 64, algo
    st = torch.cuda.current_stream().cuda_stream
    fn = lambda: abi.check(lib, lib.sb_intra_lstm_fwd(ctypes.byref(a), st), "intra")
    fn()
    torch.cuda.synchronize()
    return yf, yb, fn


def inter_call(pk, x0, x1, h, c, algo, block=1, alias=False):
    B, T, F, C = x0.shape
    y = torch.full_like(x0, float("nan"))
    hi, ci = h.clone(), c.clone()
    ho, co = (hi, ci) if alias else (torch.full_like(h, float("nan")), torch.full_like(c, float("nan")))
    a = abi.InterArgs()
    a.x0, a.x1, a.y = x0.data_ptr(), x1.data_ptr() if x1 is not None else None, y.data_ptr()
    a.h0, a.c0, a.hN, a.cN = hi.data_ptr(), ci.data_ptr(), ho.data_ptr(), co.data_ptr()
    a.dir = pk.lstm_dir(block, "inter")
    a.B, a.T, a.F, a.C, a.H, a.algo = B, T, F, C, 64, algo
    st = torch.cuda.current_stream().cuda_stream
    fn = lambda: abi.check(lib, lib.sb_inter_lstm_fwd(ctypes.byref(a), st), "inter")
    fn()
    torch.cuda.synchronize()
    return y, ho, co, fn


def timeit(fn, reps=10):
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]


if which in ("all", "parity"):
    print("inter  T=3  B=1 :", kc.check_inter(lib, "cuda:0", "dis_embed", SYN, TCP, B=1, T=3), flush=True)
    print("inter  T=40 B=2 :", kc.check_inter(lib, "cuda:0", "dis_embed", SYN, TCP, B=2, T=40, alias_state=True), flush=True)
    print("inter  T=8  B=9 :", kc.check_inter(lib, "cuda:0", "dis_embed", SYN, TCP, B=9, T=8), flush=True)
    print("intra  T=5  B=2 :", kc.check_intra(lib, "cuda:0", "dis_embed", SYN, TCP, B=2, T=5, block=1), flush=True)
    print("intra  T=300 B=1:", kc.check_intra(lib, "cuda:0", "dis_embed", SYN, TCP, B=1, T=300, block=0), flush=True)
    print("intra  T=64 B=4 :", kc.check_intra(lib, "cuda:0", "dis_embed", SYN, TCP, B=4, T=64, block=2), flush=True)
    print("tc reference    :", kc.check_intra(lib, "cuda:0", "dis_embed", SYN, TC, B=2, T=5, block=1), flush=True)

if which in ("all", "bits", "time"):
    torch.manual_seed(0)
    net = Net(**SYN).to(dev).eval()
    pk = net.engine().packed
    F, C, H = 145, 32, 64
    for (B, T) in ((32, 4), (32, 8), (5, 100), (32, 125), (32, 625)):
        g = torch.Generator(device="cpu").manual_seed(B * 1000 + T)
        x = torch.randn(B, T, F, C, generator=g).to(dev)
        x1 = torch.randn(B, T, F, C, generator=g).to(dev)
        film = torch.randn(2, B, F, C, generator=g).to(dev)
        h = (0.3 * torch.randn(B * F, H, generator=g)).to(dev)
        c = (0.3 * torch.randn(B * F, H, generator=g)).to(dev)
        rf, rb, f_old = intra_call(pk, x, film, TC)
        nf, nb, f_new = intra_call(pk, x, film, TCP)
        d = max(float((rf - nf).abs().max()), float((rb - nb).abs().max()))
        nan = bool(torch.isnan(nf).any() or torch.isnan(nb).any())
        line = "B=%3d T=%3d intra rows=%6d  maxabs(tcp - tc) = %.3e nan=%s" % (B, T, B * T, d, nan)
        if which != "bits":
            line += "   tc %9.1f us   tcp %9.1f us" % (timeit(f_old), timeit(f_new))
        print(line, flush=True)
        ry, rh, rc, f_old = inter_call(pk, x, x1, h, c, TC)
        ny, nh, nc, f_new = inter_call(pk, x, x1, h, c, TCP)
        d = max(float((ry - ny).abs().max()), float((rh - nh).abs().max()), float((rc - nc).abs().max()))
        nan = bool(torch.isnan(ny).any() or torch.isnan(nh).any() or torch.isnan(nc).any())
        line = "B=%3d T=%3d inter rows=%6d  maxabs(tcp - tc) = %.3e nan=%s" % (B, T, B * F, d, nan)
        if which != "bits":
            line += "   tc %9.1f us   tcp %9.1f us" % (timeit(f_old), timeit(f_new))
        print(line, flush=True)

if which in ("all", "golden"):
    import parity_cases as pc
    for name in ("syn_offline", "syn_nopad", "wav_syn_1m", "opi_offline"):
        print(name, pc.run_golden(lib, "cuda:0", name, TCP, TCP), flush=True)

if which == "cell7":
    # A/B of the shared-reciprocal cell update (SB_OPT_TC_CELL7): values against the one-reciprocal-per-gate form on the same
    # inputs (also with pre-activations driven far into saturation: the exponent clamps) and time per launch
    torch.manual_seed(0)
    net = Net(**SYN).to(dev).eval()
    pk = net.engine().packed
    F, C, H = 145, 32, 64

    def opt(v):
        abi.check(lib, lib.sb_set_option(abi.SB_OPT_TC_CELL7, v), "sb_set_option")
    for (B, T, scale) in ((32, 8, 1.0), (32, 32, 1.0), (32, 32, 40.0), (32, 125, 1.0), (32, 625, 1.0)):
        g = torch.Generator(device="cpu").manual_seed(B * 1000 + T)
        x = (scale * torch.randn(B, T, F, C, generator=g)).to(dev)
        x1 = torch.randn(B, T, F, C, generator=g).to(dev)
        film = (scale * torch.randn(2, B, F, C, generator=g)).to(dev)
        h = (0.3 * torch.randn(B * F, H, generator=g)).to(dev)
        c = (scale * 0.3 * torch.randn(B * F, H, generator=g)).to(dev)
        opt(0)
        rf, rb, f_old = intra_call(pk, x, film, TCP)
        t_old = timeit(f_old)
        ry, rh, rc, g_old = inter_call(pk, x, x1, h, c, TCP)
        u_old = timeit(g_old)
        opt(1)
        nf, nb, f_new = intra_call(pk, x, film, TCP)
        t_new = timeit(f_new)
        ny, nh, nc, g_new = inter_call(pk, x, x1, h, c, TCP)
        u_new = timeit(g_new)
        d_intra = max(float((rf - nf).abs().max()), float((rb - nb).abs().max()))
        d_inter = max(float((ry - ny).abs().max()), float((rh - nh).abs().max()), float((rc - nc).abs().max() / max(1.0, float(rc.abs().max()))))
        nan = bool(torch.isnan(nf).any() or torch.isnan(nb).any() or torch.isnan(ny).any() or torch.isnan(nh).any() or torch.isnan(nc).any())
        print("B=%3d T=%3d scale=%4.0f  intra maxabs %.2e  %8.1f -> %8.1f us   inter maxabs %.2e  %8.1f -> %8.1f us  nan=%s"
              % (B, T, scale, d_intra, t_old, t_new, d_inter, u_old, u_new, nan), flush=True)

if which in ("pipe", "cw16"):
    # A/B of lstm_tcr_kernel (SB_OPT_TC_PIPE: h part of the next step issued under the cell update) against lstm_tcp_kernel
    torch.manual_seed(0)
    net = Net(**SYN).to(dev).eval()
    pk = net.engine().packed
    F, C, H = 145, 32, 64

    def opt16(v):
        abi.check(lib, lib.sb_set_option(abi.SB_OPT_TC_PIPE if which == "pipe" else abi.SB_OPT_TC_CW16, v), "sb_set_option")
    for (B, T) in ((32, 8), (7, 100), (32, 32), (32, 125), (32, 625)):
        g = torch.Generator(device="cpu").manual_seed(B * 1000 + T)
        x = torch.randn(B, T, F, C, generator=g).to(dev)
        x1 = torch.randn(B, T, F, C, generator=g).to(dev)
        film = torch.randn(2, B, F, C, generator=g).to(dev)
        h = (0.3 * torch.randn(B * F, H, generator=g)).to(dev)
        c = (0.3 * torch.randn(B * F, H, generator=g)).to(dev)
        opt16(0)
        rf, rb, f_old = intra_call(pk, x, film, TCP)
        t_old = timeit(f_old)
        ry, rh, rc, g_old = inter_call(pk, x, x1 if which == "pipe" else None, h, c, TCP)
        u_old = timeit(g_old)
        opt16(1)
        nf, nb, f_new = intra_call(pk, x, film, TCP)
        t_new = timeit(f_new)
        ny, nh, nc, g_new = inter_call(pk, x, x1 if which == "pipe" else None, h, c, TCP)
        u_new = timeit(g_new)
        d_intra = max(float((rf - nf).abs().max()), float((rb - nb).abs().max()))
        d_inter = max(float((ry - ny).abs().max()), float((rh - nh).abs().max()), float((rc - nc).abs().max()))
        nan = bool(torch.isnan(nf).any() or torch.isnan(nb).any() or torch.isnan(ny).any() or torch.isnan(nh).any() or torch.isnan(nc).any())
        print("B=%3d T=%3d  intra maxabs %.2e  %8.1f -> %8.1f us   inter maxabs %.2e  %8.1f -> %8.1f us  nan=%s"
              % (B, T, d_intra, t_old, t_new, d_inter, u_old, u_new, nan), flush=True)

if which == "tcq":
    # two tiles per CTA in ping-pong (SB_ALGO_TCQ): oracle parity, bits against SB_ALGO_TCP, time per launch
    TCQ = abi.SB_ALGO_TCQ
    print("inter  T=3  B=1 :", kc.check_inter(lib, "cuda:0", "dis_embed", SYN, TCQ, B=1, T=3), flush=True)
    print("inter  T=40 B=2 :", kc.check_inter(lib, "cuda:0", "dis_embed", SYN, TCQ, B=2, T=40, alias_state=True), flush=True)
    print("inter  T=8  B=9 :", kc.check_inter(lib, "cuda:0", "dis_embed", SYN, TCQ, B=9, T=8), flush=True)
    print("intra  T=5  B=2 :", kc.check_intra(lib, "cuda:0", "dis_embed", SYN, TCQ, B=2, T=5, block=1), flush=True)
    print("intra  T=300 B=1:", kc.check_intra(lib, "cuda:0", "dis_embed", SYN, TCQ, B=1, T=300, block=0), flush=True)
    print("intra  T=64 B=5 :", kc.check_intra(lib, "cuda:0", "dis_embed", SYN, TCQ, B=5, T=64, block=2), flush=True)
    torch.manual_seed(0)
    net = Net(**SYN).to(dev).eval()
    pk = net.engine().packed
    F, C, H = 145, 32, 64
    for (B, T) in ((32, 8), (32, 32), (7, 100), (32, 125), (32, 625)):
        g = torch.Generator(device="cpu").manual_seed(B * 1000 + T)
        x = torch.randn(B, T, F, C, generator=g).to(dev)
        x1 = torch.randn(B, T, F, C, generator=g).to(dev)
        film = torch.randn(2, B, F, C, generator=g).to(dev)
        h = (0.3 * torch.randn(B * F, H, generator=g)).to(dev)
        c = (0.3 * torch.randn(B * F, H, generator=g)).to(dev)
        rf, rb, f_old = intra_call(pk, x, film, TCP)
        nf, nb, f_new = intra_call(pk, x, film, TCQ)
        d = max(float((rf - nf).abs().max()), float((rb - nb).abs().max()))
        nan = bool(torch.isnan(nf).any() or torch.isnan(nb).any())
        print("B=%3d T=%3d intra rows=%6d  maxabs(tcq - tcp) = %.3e nan=%s   tcp %9.1f us   tcq %9.1f us"
              % (B, T, B * T, d, nan, timeit(f_old), timeit(f_new)), flush=True)
        ry, rh, rc, g_old = inter_call(pk, x, x1, h, c, TCP)
        ny, nh, nc, g_new = inter_call(pk, x, x1, h, c, TCQ)
        d = max(float((ry - ny).abs().max()), float((rh - nh).abs().max()), float((rc - nc).abs().max()))
        nan = bool(torch.isnan(ny).any() or torch.isnan(nh).any() or torch.isnan(nc).any())
        print("B=%3d T=%3d inter rows=%6d  maxabs(tcq - tcp) = %.3e nan=%s   tcp %9.1f us   tcq %9.1f us"
              % (B, T, B * F, d, nan, timeit(g_old), timeit(g_new)), flush=True)
