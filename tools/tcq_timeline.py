#!/usr/bin/env python
"""Timeline of one lstm_tcq_kernel CTA (debug build: NVCC flag -DSB_TCQ_DEBUG): globaltimer stamps of the stream group's and
the cell group's phases for the first steps, printed relative to the first event."""
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
B, T, F, C = 32, 8, 145, 32
x = torch.randn(B, T, F, C, device=dev); film = torch.randn(2, B, F, C, device=dev)
yf, yb = torch.empty_like(x), torch.empty_like(x)
a = abi.IntraArgs()
a.x, a.y_fwd, a.y_bwd = x.data_ptr(), yf.data_ptr(), yb.data_ptr()
a.film_scale, a.film_shift = film[0].data_ptr(), film[1].data_ptr()
a.dir[0], a.dir[1] = pk.lstm_dir(1, "intra0"), pk.lstm_dir(1, "intra1")
WHICH = sys.argv[1] if len(sys.argv) > 1 else "tcq"          # "tcq": lstm_tcq_kernel, "tcr": lstm_tcr_kernel (SB_ALGO_TC, the default route)
a.B, a.T, a.F, a.C, a.H, a.algo = B, T, F, C, 64, abi.SB_ALGO_TCQ if WHICH == "tcq" else abi.SB_ALGO_TC
st = torch.cuda.current_stream().cuda_stream
fn = lib._cdll.sb_tcq_debug_read if hasattr(lib, "_cdll") else ctypes.CDLL(_lib.LIB_PATH).sb_tcq_debug_read
buf = (ctypes.c_longlong * (4 * 4096))()
for rep in range(2):
    abi.check(lib, lib.sb_intra_lstm_fwd(ctypes.byref(a), st), "intra"); torch.cuda.synchronize()
    n = fn(buf, 4096)
ev = sorted((buf[4 * i], buf[4 * i + 1], buf[4 * i + 2], buf[4 * i + 3]) for i in range(n))
t0 = ev[0][0]
names = {100: "S build.begin", 101: "S build.done", 102: "S hready.seen", 103: "S proj.issued", 104: "S proj.read", 105: "S gates.issued", 106: "S iter.end",
         200: "C wait.begin", 201: "C gates.seen", 202: "C cell.done",
         300: "M wait.h", 301: "M h.seen", 302: "M proj.issued", 303: "M gates.begin", 304: "M gates.issued",
         400: "S iter.begin", 401: "S pdone.seen", 402: "S emit.done", 403: "S bar.done", 404: "S prepare.done",
         500: "C gates.seen", 501: "C chunk0.done", 502: "C chunk1.done", 503: "C chunk2.done", 504: "C chunk3.done",
         600: "M xready.seen", 601: "M xpart.issued", 610: "M hk0.seen", 611: "M hk1.seen", 612: "M hk2.seen", 613: "M hk3.seen",
         630: "M iter.issued", 510: "C first.ld.done"}
for _c in range(4):
    names[511 + 3 * _c] = "C chunk%d.computed" % _c
    names[512 + 3 * _c] = "C chunk%d.stored" % _c
for t, code, X, sw in ev:
    step, warp = sw // 100, sw % 100
    if step < 2 or step > 6: continue
    if code // 100 == 4 and warp != 0: continue
    if code // 100 == 5 and warp not in (4, 8): continue
    if code // 100 == 1 and warp not in (0, 1): continue
    if code // 100 == 2 and warp not in (4, 8): continue
    if code // 100 == 3 and warp != 12: continue
    print("%8.2f us  %-16s tile %d step %d warp %d" % ((t - t0) / 1e3, names.get(code, code), X, step, warp))
