import sys, torch
sys.path.insert(0, "/root/repo")
from oetr_b200 import superglue as sg
g = torch.Generator().manual_seed(0)
n = 2048
q, k, v = (torch.randn(1, 64, 4, n, generator=g).cuda() for _ in range(3))
s = (torch.randn(1, n, n, generator=g) * 2).cuda()
for it in (1, 10, 100):
    sg.log_optimal_transport(s, 1.0, it)
sg.attention(q, k, v); sg.attention(q, k, v, mode="fp32")
torch.cuda.synchronize()
