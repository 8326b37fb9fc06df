import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tokenpacker_b200 import TokenPackerB200
m = TokenPackerB200(hidden_size=4096, scale_factor=2).to("cuda", torch.bfloat16).train()
n = 64
x0 = torch.randn(n, 576, 1024, device="cuda").bfloat16()
xm = torch.randn(n, 576, 4096, device="cuda").bfloat16()
gw = torch.randn(n, 144, 4096, device="cuda").bfloat16()
for _ in range(2):
    for p in m.parameters():
        p.grad = None
    out = m((x0, xm))
    out.backward(gw)
torch.cuda.synchronize()
