"""Diagnose tests/test_projector_gpu.py::test_kernel_variants_agree_bitwise: which mode / rows / columns differ, and whether
each configuration is run-to-run deterministic."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tokenpacker_b200.kernels import gemm_bf16
from tokenpacker_b200 import TokenPackerB200

torch.manual_seed(3)
s, hidden, n = 4, 256, 10
m = TokenPackerB200(hidden_size=hidden, scale_factor=s).cuda().bfloat16()
with torch.no_grad():
    for name, p in m.named_parameters():
        if p.dim() == 1:
            p.add_(0.1 * torch.randn_like(p))
g = torch.Generator(device="cuda").manual_seed(7)
x0 = torch.randn(n, 576, 1024, device="cuda", generator=g).bfloat16()
xm = torch.randn(n, 576, 4096, device="cuda", generator=g).bfloat16()
outs = {}
with torch.no_grad():
    for mode in ("0", "1", "2", "3"):
        os.environ["TP_GEMM_MODE"] = mode
        for rep in range(3):
            outs[(mode, "full", rep)] = m((x0, xm)).clone()
            outs[(mode, "half", rep)] = m((x0[5:], xm[5:])).clone()
        torch.cuda.synchronize()


def diff(a, b, tag):
    d = (a.float() - b.float()).abs()
    nd = int((d > 0).sum())
    msg = f"{tag}: n_diff={nd} of {d.numel()} max={d.max().item():.3e}"
    if nd:
        idx = torch.nonzero(d > 0)
        rows = sorted(set((int(i[0]), int(i[1])) for i in idx))[:6]
        cols = sorted(set(int(i[2]) for i in idx))[:12]
        msg += f" first (crop,token)={rows} cols={cols}"
    print(msg)


for mode in ("0", "1", "2", "3"):
    for kind in ("full", "half"):
        diff(outs[(mode, kind, 0)], outs[(mode, kind, 1)], f"mode {mode} {kind} rep0 vs rep1")
        diff(outs[(mode, kind, 0)], outs[(mode, kind, 2)], f"mode {mode} {kind} rep0 vs rep2")
    diff(outs[(mode, "full", 0)][5:], outs[(mode, "half", 0)], f"mode {mode} full[5:] vs half")
    diff(outs[("0", "full", 0)], outs[(mode, "full", 0)], f"mode 0 vs mode {mode} full")

# GEMM alone: identical A rows at different positions / problem sizes
torch.manual_seed(0)
for (mm, nn, kk, gelu) in [(360, 256, 1024, True), (360, 1024, 1024, False), (2880, 1024, 4096, True)]:
    a = torch.randn(mm, kk, device="cuda").bfloat16()
    a[mm // 2:] = a[:mm - mm // 2]
    b = (torch.randn(nn, kk, device="cuda") * 0.03).bfloat16()
    bias = torch.randn(nn, device="cuda")
    for mode in ("1", "2", "3"):
        os.environ["TP_GEMM_MODE"] = mode
        c = gemm_bf16(a, b, bias=bias, gelu=gelu).clone()
        c2 = gemm_bf16(a, b, bias=bias, gelu=gelu).clone()
        chalf = gemm_bf16(a[:mm // 2].contiguous(), b, bias=bias, gelu=gelu).clone()
        diff(c, c2, f"gemm M={mm} N={nn} K={kk} gelu={gelu} mode {mode} rerun")
        diff(c[mm // 2:], c[:mm - mm // 2], f"   rows shifted by {mm // 2}")
        diff(c[:mm // 2], chalf, f"   first half vs half-size problem")
