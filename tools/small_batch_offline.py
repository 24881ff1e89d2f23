import sys, torch
sys.path.insert(0, ".")
from bench import SYN, radius_one_hot, synthetic_clips
from sound_bubble_b200 import Net
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = Net(**SYN).to(dev).eval()
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n
for B in (12, 16, 24):
    x = synthetic_clips(B, 1).to(dev); dis = radius_one_hot(B).to(dev)
    inp = {"mixture": x, "dis_embed": dis}
    net.pipeline_offline = False
    t0 = timed(lambda: net(inp))
    ref = net(inp)["output"]
    net.pipeline_offline = True
    res = []
    for tc in (125,):
        net.offline_min_rows, net.offline_slice_frames = 0, tc
        for ia in (None, 7):
            net.offline_inter_algo = ia
            t1 = timed(lambda: net(inp))
            err = float((net(inp)["output"] - ref).abs().max())
            res.append("slice %d inter %s: %.2f ms (%.0e)" % (tc, ia, t1, err))
    print("B=%d single call %.2f ms | " % (B, t0) + " | ".join(res), flush=True)
