import sys, torch
sys.path.insert(0, ".")
from bench import SYN, radius_one_hot
from sound_bubble_b200 import Net
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = Net(**dict(SYN, use_attn=True)).to(dev).eval()
net.pipeline_offline = False
B = 32
dis = radius_one_hot(B).to(dev)
g = torch.Generator().manual_seed(1)
x = (0.1 * torch.randn(B, 6, 192 * 125 + 96, generator=g)).to(dev)
for _ in range(2):
    net({"mixture": x, "dis_embed": dis}, pad=False)
torch.cuda.synchronize()
