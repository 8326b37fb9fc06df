import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tokenpacker_b200 import TokenPackerB200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
m = TokenPackerB200(hidden_size=4096, scale_factor=2).to("cuda", torch.bfloat16).eval()
x0 = torch.randn(n, 576, 1024, device="cuda").bfloat16()
xm = torch.randn(n, 576, 4096, device="cuda").bfloat16()
with torch.no_grad():
    for _ in range(4):
        m((x0, xm))
torch.cuda.synchronize()
