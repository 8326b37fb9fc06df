import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tokenpacker_b200 import TokenPackerB200, hd_assemble
from tokenpacker_b200 import synthetic as syn
s, hidden, n = 4, 256, 10
grids = [(2, 2), (1, 1), (1, 3)]
m = TokenPackerB200(hidden_size=hidden, scale_factor=s)
m.load_state_dict({k: torch.from_numpy(v) for k, v in syn.synthetic_state_dict(hidden, seed=3).items()})
m = m.to("cuda", torch.bfloat16).eval()
g = torch.Generator(device="cuda").manual_seed(7)
x0 = torch.randn(n, 576, 1024, device="cuda", generator=g).bfloat16()
xm = torch.randn(n, 576, 4096, device="cuda", generator=g).bfloat16()
sep = torch.randn(hidden, device="cuda", generator=g).bfloat16()
ret = torch.randn(hidden, device="cuda", generator=g).bfloat16()
hb, wb = [a for a, _ in grids], [b for _, b in grids]
with torch.no_grad():
    full = m((x0, xm))
    for rep in range(4):
        again = m((x0, xm))
        print('rerun equal', torch.equal(full, again), int(((full.float()-again.float()).abs().amax(-1) > 0).sum()), 'rows differ')
    a1 = m((x0[:5], xm[:5])); a2 = m((x0[:5], xm[:5])); print('half rerun equal', torch.equal(a1, a2))
    a = m((x0[:5], xm[:5])); b = m((x0[5:], xm[5:]))
    print("first half equal", torch.equal(full[:5], a), "second half equal", torch.equal(full[5:], b))
    for i in range(10):
        loc = a[i] if i < 5 else b[i - 5]
        d = (full[i].float() - loc.float()).abs()
        if d.max() > 0:
            print("crop", i, "max diff", d.max().item(), "n diff", int((d > 0).sum()), "rows", torch.nonzero(d.amax(-1) > 0).flatten()[:10].tolist())
    packed, cu = m.forward_packed((x0, xm), hb, wb, sep, ret)
    p2, cu2 = hd_assemble(full, hb, wb, sep, ret)
    print("packed equal", torch.equal(packed, p2), (packed.float() - p2.float()).abs().max().item())
    if not torch.equal(packed, p2):
        d = (packed.float() - p2.float()).abs().amax(-1)
        print("rows differing", torch.nonzero(d > 0).flatten()[:20].tolist(), "of", packed.shape[0])
