import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from tokenpacker_b200 import TokenPackerB200
m = TokenPackerB200(hidden_size=4096, scale_factor=2).to("cuda", torch.bfloat16).eval()
n = 64
x0 = torch.randn(n, 576, 1024).bfloat16().pin_memory(); xm = torch.randn(n, 576, 4096).bfloat16().pin_memory()
out = torch.empty(n, 144, 4096, dtype=torch.bfloat16).pin_memory()
with torch.no_grad():
    ref = m((x0.cuda(), xm.cuda())).cpu()
    for chunk in (8, 16, 4, 8):
        for _ in range(3): m.forward_host((x0, xm), out=out, chunk_crops=chunk)
        t0 = time.perf_counter()
        for _ in range(30): m.forward_host((x0, xm), out=out, chunk_crops=chunk)
        dt = (time.perf_counter() - t0) / 30
        print(f"chunk {chunk}: {dt*1e3:.3f} ms/step  {n*144/dt/1e6:.3f} M tok/s  bits_equal={torch.equal(out, ref)}")
