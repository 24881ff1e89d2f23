#!/usr/bin/env python
"""Times the intra / inter LSTM entry points of the C ABI per kernel family at streaming and offline shapes
(CUDA events, L2 flushed between launches).  Used to calibrate pick_algo in sb_lstm.cu."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import SYN  # noqa: E402
from sound_bubble_b200 import Net, _abi as abi, _lib  # noqa: E402

ALGO = {1: "tile", 2: "lane1", 3: "lane2", 4: "lane4", 5: "ws", 8: "ws2", 7: "tc"}


def main():
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    if len(sys.argv) > 1 and sys.argv[1] == "pdl":
        _lib.set_pdl(True)
    only_tc = len(sys.argv) > 1 and sys.argv[1] == "tc"
    warm = len(sys.argv) > 1 and sys.argv[1] == "warm"       # no L2 flush: weights and state stay L2-resident (steady state)
    torch.manual_seed(0)
    net = Net(**SYN).to(dev).eval()
    pk = net.engine().packed
    F, C, H = 145, 32, 64
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def timeit(fn, reps):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            if not warm:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    for (B, T) in (((32, 1), (8, 1)) if warm else ((4, 625), (32, 625)) if only_tc else ((32, 1), (8, 1), (128, 1), (32, 8), (4, 625), (32, 625))):
        x = torch.randn(B, T, F, C, device=dev)
        y0, y1 = torch.empty_like(x), torch.empty_like(x)
        h = torch.zeros(B * F, H, device=dev); c = torch.zeros(B * F, H, device=dev)
        film = torch.randn(2, B, F, C, device=dev)
        for kind in ("intra", "inter"):
            for algo in ((1, 7) if only_tc else (1, 2, 3, 4, 5, 8, 7)):
                rows = B * T if kind == "intra" else B * F
                steps = F if kind == "intra" else T
                ctas = {1: rows / 8 / 8, 2: rows, 3: rows / 2, 4: rows / 4, 5: rows / 3, 8: rows / 6, 7: rows / 128}[algo] * (2 if kind == "intra" else 1)
                if algo != 1 and ctas * steps > 148 * 145 * 80:
                    continue                                    # hopeless: skip the very long lane runs
                if kind == "intra":
                    a = abi.IntraArgs()
                    a.x, a.y_fwd, a.y_bwd = x.data_ptr(), y0.data_ptr(), y1.data_ptr()
                    a.film_scale, a.film_shift = film[0].data_ptr(), film[1].data_ptr()
                    a.dir[0], a.dir[1] = pk.lstm_dir(1, "intra0"), pk.lstm_dir(1, "intra1")
                    a.B, a.T, a.F, a.C, a.H, a.algo = B, T, F, C, H, algo
                    fn = lambda: abi.check(lib, lib.sb_intra_lstm_fwd(ctypes.byref(a), st), "intra")
                else:
                    a = abi.InterArgs()
                    a.x0, a.x1, a.y = x.data_ptr(), y0.data_ptr(), y1.data_ptr()
                    a.h0, a.c0, a.hN, a.cN = h.data_ptr(), c.data_ptr(), h.data_ptr(), c.data_ptr()
                    a.dir = pk.lstm_dir(1, "inter")
                    a.B, a.T, a.F, a.C, a.H, a.algo = B, T, F, C, H, algo
                    fn = lambda: abi.check(lib, lib.sb_inter_lstm_fwd(ctypes.byref(a), st), "inter")
                us = timeit(fn, 5 if T > 8 else 20)
                gflop = 2.0 * rows * steps * (4 * H * (C + H) + H * C) * (2 if kind == "intra" else 1) / 1e9
                print("B=%3d T=%3d %-5s %-5s rows=%6d steps=%3d  %10.1f us  %7.2f TFLOP/s  %8.1f ns/step"
                      % (B, T, kind, ALGO[algo], rows, steps, us, gflop / us * 1e-3 * 1e3 / 1e3 * 1e3 / 1e3 if False else gflop / (us * 1e-6) / 1e3,
                         us * 1e3 / steps), flush=True)


if __name__ == "__main__":
    main()
