"""What does cuBLAS launch for our GEMM shapes?  (ncu --set launchstats style probe; run under ncu)"""
import sys
import torch
m, n, k = (int(v) for v in sys.argv[1:4])
a = torch.randn(m, k, device="cuda").bfloat16()
b = torch.randn(n, k, device="cuda").bfloat16()
for _ in range(3):
    c = torch.matmul(a, b.t())
torch.cuda.synchronize()
