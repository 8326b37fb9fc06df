import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import tokenpacker_oracle as tpo
from oracle import torch_port
from tokenpacker_b200 import TokenPackerB200
for s, hidden, n in [(2, 256, 2), (3, 128, 3), (4, 256, 4)]:
    params = {k: tpo.round_bf16(v) for k, v in tpo.make_params(hidden, seed=21 + s).items()}
    m = TokenPackerB200(hidden_size=hidden, scale_factor=s)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    m = m.to("cuda", torch.bfloat16).train()
    x0, xm = tpo.make_inputs(n, seed=31 + s)
    x0 = torch.from_numpy(tpo.round_bf16(x0)).cuda(); xm = torch.from_numpy(tpo.round_bf16(xm)).cuda()
    gw = torch.randn(n, (24 // s) ** 2, hidden, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    out = m((x0.bfloat16(), xm.bfloat16()))
    (out.float() * gw.bfloat16().float()).sum().backward()
    ref_p = {k: torch.from_numpy(v).cuda().requires_grad_(True) for k, v in params.items()}
    ref_out = torch_port.forward(ref_p, x0, xm, s)
    (ref_out * gw.bfloat16().float()).sum().backward()
    print(f"--- s={s} H={hidden} n={n}")
    for name, p in m.named_parameters():
        g, r = p.grad.float(), ref_p[name].grad
        if name == "clip_attn.in_proj_bias":
            for i, nm in enumerate("qkv"):
                gg, rr = g[i*1024:(i+1)*1024], r[i*1024:(i+1)*1024]
                print(f"{name+'['+nm+']':32s} ref_rms {rr.pow(2).mean().sqrt().item():.3e} err_rms {(gg-rr).pow(2).mean().sqrt().item():.3e}")
        else:
            print(f"{name:32s} ref_rms {r.pow(2).mean().sqrt().item():.3e} err_rms {(g-r).pow(2).mean().sqrt().item():.3e} rel {(g-r).pow(2).mean().sqrt().item()/(r.pow(2).mean().sqrt().item()+1e-30):.3e}")
