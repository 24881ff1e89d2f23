import sys, torch
sys.path.insert(0, ".")
from bench import SYN, radius_one_hot
from sound_bubble_b200 import Net
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = Net(**dict(SYN, use_attn=True)).to(dev).eval()
B = 32
dis = radius_one_hot(B).to(dev)
g = torch.Generator().manual_seed(1)
x = (0.1 * torch.randn(B, 6, 192 * 6 + 96, generator=g)).to(dev)
sess = net.streaming(B, dis, use_graph=False)
for t in range(6):
    sess.feed(x[..., t * 192: t * 192 + 288])
torch.cuda.synchronize()
