import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tokenpacker_b200 import TokenPackerB200
from tokenpacker_b200 import synthetic as syn
s, hidden, n = 4, 256, 10
m = TokenPackerB200(hidden_size=hidden, scale_factor=s)
m.load_state_dict({k: torch.from_numpy(v) for k, v in syn.synthetic_state_dict(hidden, seed=3).items()})
m = m.to("cuda", torch.bfloat16).eval()
g = torch.Generator(device="cuda").manual_seed(7)
x0 = torch.randn(n, 576, 1024, device="cuda", generator=g).bfloat16()
xm = torch.randn(n, 576, 4096, device="cuda", generator=g).bfloat16()
res = {}
with torch.no_grad():
    for mode in ("1", "2", "3"):
        os.environ["TP_GEMM_MODE"] = mode
        res["full" + mode] = m((x0, xm)).clone()
        res["a" + mode] = m((x0[:5], xm[:5])).clone()
        res["one" + mode] = m((x0[2:3], xm[2:3])).clone()
def cmp(x, y, sl=None):
    a, b = res[x], res[y]
    if sl is not None: a = a[sl]
    d = (a.float() - b.float()).abs().amax(-1)
    return f"{x} vs {y}: {'EQUAL' if (d == 0).all() else 'rows differ: ' + str(torch.nonzero(d > 0).tolist()[:6])}"
print(cmp("full1", "full2")); print(cmp("full1", "full3")); print(cmp("a1", "a2")); print(cmp("a1", "a3"))
for mode in ("1", "2", "3"):
    print(cmp("full" + mode, "a" + mode, slice(0, 5)))
    print(cmp("full" + mode, "one" + mode, slice(2, 3)))
