#!/bin/bash
mkdir -p gpurun_out
timeout 600 python - <<'PY'
import torch, sys, time
sys.path.insert(0, ".")
from bench import SYN, synthetic_clips, radius_one_hot
from sound_bubble_b200 import Net
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = Net(**SYN).to(dev).eval()
x = synthetic_clips(32, 1234).to(dev); dis = radius_one_hot(32).to(dev)
inp = {"mixture": x, "dis_embed": dis}
net.pipeline_offline = False
ref = net(inp)["output"]
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n
print("single call %.2f ms" % t(lambda: net(inp)))
net.pipeline_offline = True
for inter, tc in ((7, 250), (7, 157), (7, 125), (7, 105), (7, 79), (7, 63), (1, 125), (5, 63)):
    net.offline_inter_algo = inter
    net.offline_slice_frames = tc
    y = net(inp)["output"]
    err = float((y - ref).abs().max())
    print("inter algo %d slice %3d frames: %.2f ms  maxabs vs single %.2e" % (inter, tc, t(lambda: net(inp)), err), flush=True)
PY
