"""Run one GEMM shape a few times (for `ncu -k regex:tp_gemm -s 3 -c 1 python tools/prof_shape.py M N K [gelu] [mode]`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
m, n, k = (int(v) for v in sys.argv[1:4])
gelu = len(sys.argv) > 4 and sys.argv[4] == "1"
if len(sys.argv) > 5:
    os.environ["TP_GEMM_MODE"] = sys.argv[5]
import torch  # noqa: E402
from tokenpacker_b200.kernels import gemm_bf16  # noqa: E402

a = torch.randn(m, k, device="cuda").bfloat16()
b = (torch.randn(n, k, device="cuda") * 0.02).bfloat16()
bias = torch.randn(n, device="cuda")
c = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
for _ in range(6):
    gemm_bf16(a, b, bias=bias, gelu=gelu, out=c)
torch.cuda.synchronize()
