"""Which clipped-box TMA store forms does the hardware accept?  Each case runs in its own process (a device-side fault kills only
that case) and compares the packed / fused forward with the separate-kernel plan.

    python tools/tma_store_probe.py            # on a B200
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASE = r'''
import os, sys
sys.path.insert(0, %(root)r)
import torch
from tokenpacker_b200 import TokenPackerB200, hd_assemble
from tokenpacker_b200 import synthetic as syn
s, hidden, kind = %(s)d, %(hidden)d, %(kind)r
m = TokenPackerB200(hidden_size=hidden, scale_factor=s)
m.load_state_dict({k: torch.from_numpy(v) for k, v in syn.synthetic_state_dict(hidden, seed=3).items()})
m = m.to("cuda", torch.bfloat16).eval()
g = torch.Generator(device="cuda").manual_seed(1)
grids = [(1, 1), (2, 3), (3, 1), (1, 2), (2, 2)]
n = sum(a * b + (1 if a * b > 1 else 0) for a, b in grids)
x0 = torch.randn(n, 576, 1024, device="cuda", generator=g).bfloat16()
xm = torch.randn(n, 576, 4096, device="cuda", generator=g).bfloat16()
sep = torch.randn(hidden, device="cuda", generator=g).bfloat16()
ret = torch.randn(hidden, device="cuda", generator=g).bfloat16()
hb, wb = [a for a, _ in grids], [b for _, b in grids]
with torch.no_grad():
    if kind == "packed":
        out, _ = m.forward_packed((x0, xm), hb, wb, sep, ret)
        torch.cuda.synchronize()
        ref, _ = hd_assemble(m((x0, xm)), hb, wb, sep, ret)
        torch.cuda.synchronize()
        print("RESULT", "bit-exact" if torch.equal(out, ref) else "MISMATCH rows %%d" %% int((out != ref).any(-1).sum()))
    else:
        out = m((x0, xm))
        torch.cuda.synchronize()
        os.environ["TP_FUSE_ATTN"] = "0"
        ref = m((x0, xm))
        torch.cuda.synchronize()
        d = (out.float() - ref.float())
        print("RESULT rel-rms vs separate kernels %%.3e" %% float(d.pow(2).mean().sqrt() / ref.float().pow(2).mean().sqrt()))
'''

cases = [("packed s=3 box 64 rows, aligned, unclipped", dict(s=3, hidden=256, kind="packed"), {"TP_FUSE_ATTN": "0", "TP_GEMM_MODE": "2"}),
         ("packed s=2 box 128 rows, clipped", dict(s=2, hidden=512, kind="packed"), {"TP_FUSE_ATTN": "0", "TP_GEMM_MODE": "2"}),
         ("packed s=4 box 36 rows, 512B-aligned sources", dict(s=4, hidden=512, kind="packed"), {"TP_FUSE_ATTN": "0", "TP_GEMM_MODE": "2"}),
         ("packed s=4, no swizzle", dict(s=4, hidden=512, kind="packed"), {"TP_FUSE_ATTN": "0", "TP_GEMM_MODE": "2", "TP_SEG_NOSWIZZLE": "1"}),
         ("packed s=2, no swizzle", dict(s=2, hidden=512, kind="packed"), {"TP_FUSE_ATTN": "0", "TP_GEMM_MODE": "2", "TP_SEG_NOSWIZZLE": "1"}),
         ("fused s=2 (5-D window-major stores)", dict(s=2, hidden=256, kind="fused"), {}),
         ("fused s=4 (5-D window-major stores)", dict(s=4, hidden=256, kind="fused"), {}),
         ("fused s=2, no swizzle", dict(s=2, hidden=256, kind="fused"), {"TP_SEG_NOSWIZZLE": "1"}),
         ("fused s=4, no swizzle", dict(s=4, hidden=256, kind="fused"), {"TP_SEG_NOSWIZZLE": "1"}),
         ("fused s=2 packed output", dict(s=2, hidden=512, kind="packed"), {}),
         ("fused s=4 packed output, no swizzle", dict(s=4, hidden=512, kind="packed"), {"TP_SEG_NOSWIZZLE": "1"})]
for name, kw, env in cases:
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", CASE % dict(root=ROOT, **kw)], capture_output=True, text=True, env=e, timeout=300)
    res = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    err = [l for l in (r.stderr + r.stdout).splitlines() if "rror" in l or "timed out" in l]
    print(f"{name:55s} -> {res[0] if res else 'FAILED: ' + (err[-1][:160] if err else 'rc %d' % r.returncode)}", flush=True)
