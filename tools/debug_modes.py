import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tokenpacker_b200.kernels import gemm_bf16
torch.manual_seed(0)
for (m, n, k, gelu) in [(360, 1024, 1024, False), (360, 256, 1024, True), (512, 1024, 4096, False), (360, 256, 256, False)]:
    a = torch.randn(m, k, device="cuda").bfloat16()
    b = (torch.randn(n, k, device="cuda") * 0.03).bfloat16()
    bias = torch.randn(n, device="cuda")
    outs = {}
    for mode in ("1", "2"):
        os.environ["TP_GEMM_MODE"] = mode
        outs[mode] = gemm_bf16(a, b, bias=bias, gelu=gelu).clone()
    d = (outs["1"].float() - outs["2"].float()).abs()
    ref = torch.nn.functional.gelu(a.float() @ b.float().t() + bias) if gelu else a.float() @ b.float().t() + bias
    print(f"M={m} N={n} K={k} gelu={gelu}: n_diff={int((d > 0).sum())} of {d.numel()} max={d.max().item():.3e}; "
          f"err1={(outs['1'].float()-ref).abs().max().item():.3e} err2={(outs['2'].float()-ref).abs().max().item():.3e}")
    if (d > 0).any():
        idx = torch.nonzero(d > 0)[:5]
        for r, c in idx.tolist():
            print("   ", r, c, outs["1"][r, c].item(), outs["2"][r, c].item(), ref[r, c].item())
