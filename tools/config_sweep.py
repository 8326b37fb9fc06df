"""Throughput of every BASELINE.json configuration that fits one GPU (device-resident inputs, CUDA events)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from tokenpacker_b200 import TokenPackerB200, hd_grid  # noqa: E402
from tokenpacker_b200 import synthetic as syn  # noqa: E402
from tokenpacker_b200.hd import n_crops  # noqa: E402

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops_sustained": 1400.0}
PEAK = peaks["bf16_tflops_sustained"]


def timed(fn, warm=5, iters=30):
    with torch.no_grad():
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def module(s):
    m = TokenPackerB200(hidden_size=4096, scale_factor=s)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in syn.synthetic_state_dict(4096, seed=0).items()})
    return m.to("cuda", torch.bfloat16).eval()


rows = []
for name, n, s in [("configs[1] N=64 s=2", 64, 2), ("configs[2] N=128 s=2", 128, 2), ("configs[2] N=128 s=3", 128, 3),
                   ("configs[2] N=128 s=4", 128, 4), ("latency N=1 s=2", 1, 2), ("latency N=8 s=2", 8, 2)]:
    m = module(s)
    x0 = torch.randn(n, 576, 1024, device="cuda").bfloat16()
    xm = torch.randn(n, 576, 4096, device="cuda").bfloat16()
    ms = timed(lambda: m((x0, xm)))
    tok = n * (24 // s) ** 2
    tf = syn.flops_per_crop(s) * n / ms / 1e9
    rows.append({"config": name, "ms": round(ms, 4), "tokens_per_s": round(tok / ms * 1e3), "tflops_alg": round(tf, 1), "frac_sustained_peak": round(tf / PEAK, 3)})
    print(rows[-1], flush=True)
    del m, x0, xm

# serving latency with the forward captured in a CUDA graph (removes Python / launch overhead)
for n in (1, 2, 4, 10, 16):
    m = module(2)
    sx0 = torch.randn(n, 576, 1024, device="cuda").bfloat16()
    sxm = torch.randn(n, 576, 4096, device="cuda").bfloat16()
    with torch.no_grad():
        m((sx0, sxm))
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            so = m((sx0, sxm))
    ms = timed(lambda: gr.replay(), warm=10, iters=100)
    ms_eager = timed(lambda: m((sx0, sxm)), warm=10, iters=100)
    rows.append({"config": f"latency N={n} s=2, CUDA graph replay", "ms": round(ms, 4), "ms_eager_python": round(ms_eager, 4),
                 "tokens_per_s": round(n * 144 / ms * 1e3)})
    print(rows[-1], flush=True)
    del m, gr

# configs[3]: HD patch_num=9, s=2, 32 images, packed output — (a) seeded image sizes -> grids via the grid selector (231 crops),
# (b) the all-1088x1088 variant: 32 x (3x3 + thumbnail) = 320 crops, 1450 tokens per image
g = torch.Generator().manual_seed(0)
hs = torch.randint(224, 1345, (32,), generator=g).tolist()
ws_ = torch.randint(224, 1345, (32,), generator=g).tolist()
for label, sizes in (("seeded sizes", list(zip(hs, ws_))), ("all 1088x1088", [(1088, 1088)] * 32)):
    grids = [hd_grid(h, w, 9) for h, w in sizes]
    n = sum(n_crops(a, b) for a, b in grids)
    m = module(2)
    x0 = torch.randn(n, 576, 1024, device="cuda").bfloat16()
    xm = torch.randn(n, 576, 4096, device="cuda").bfloat16()
    sep = torch.randn(4096, device="cuda").bfloat16()
    ret = torch.randn(4096, device="cuda").bfloat16()
    hb, wb = [a for a, _ in grids], [b for _, b in grids]
    ms = timed(lambda: m.forward_packed((x0, xm), hb, wb, sep, ret))
    with torch.no_grad():
        _, cu = m.forward_packed((x0, xm), hb, wb, sep, ret)
    tf = syn.flops_per_crop(2) * n / ms / 1e9
    rows.append({"config": f"configs[3] HD patch_num=9 s=2, 32 images ({label}) -> {n} crops, packed {int(cu[-1])} rows ({int(cu[-1]) / 32:.0f} tok/img)",
                 "ms": round(ms, 4), "tokens_per_s": round(n * 144 / ms * 1e3), "tflops_alg": round(tf, 1), "frac_sustained_peak": round(tf / PEAK, 3)})
    print(rows[-1], flush=True)
    del m, x0, xm
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "config_sweep.json"), "w"), indent=1)
