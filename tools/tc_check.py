#!/usr/bin/env python
"""Stand-alone parity + timing check of the tcgen05 LSTM kernel (SB_ALGO_TC), run in its own process under `timeout`
so that a trap in an experimental kernel cannot take the rest of a test session with it."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kernel_cases as kc  # noqa: E402
from oracle.cases import SYN  # noqa: E402
from sound_bubble_b200 import _abi as abi, _lib  # noqa: E402

lib = _lib.load()
TC = abi.SB_ALGO_TC
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "parity"):
    print("inter  T=3  B=1 :", kc.check_inter(lib, "cuda:0", "dis_embed", SYN, TC, B=1, T=3), flush=True)
    print("inter  T=40 B=2 :", kc.check_inter(lib, "cuda:0", "dis_embed", SYN, TC, B=2, T=40, alias_state=True), flush=True)
    print("intra  T=5  B=2 :", kc.check_intra(lib, "cuda:0", "dis_embed", SYN, TC, B=2, T=5, block=1), flush=True)
    print("intra  T=300 B=1:", kc.check_intra(lib, "cuda:0", "dis_embed", SYN, TC, B=1, T=300, block=0), flush=True)
    print("tile reference  :", kc.check_intra(lib, "cuda:0", "dis_embed", SYN, abi.SB_ALGO_TILE, B=2, T=5, block=1), flush=True)
if which in ("all", "golden"):
    import parity_cases as pc
    for name in ("syn_offline", "syn_nopad", "wav_syn_1m", "opi_offline"):
        print(name, pc.run_golden(lib, "cuda:0", name, TC, TC), flush=True)
