"""Stage-by-stage bring-up on a real B200 (each stage in its own process so a trap cannot poison the next).
    gpurun -- python tools/first_light.py            -> gpurun_out/first_light.log
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = {}


def stage(f):
    STAGES[f.__name__] = f
    return f


def _gemm_case(m, n, k, **kw):
    import torch
    from tokenpacker_b200.kernels import gemm_bf16
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(m, k, device="cuda", generator=g).bfloat16()
    b = (torch.randn(n, k, device="cuda", generator=g) * 0.05).bfloat16()
    out = gemm_bf16(a, b, **kw)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    err = (out.float() - ref).abs().max().item()
    print(f"gemm {m}x{n}x{k}: max err {err:.4e} (ref max {ref.abs().max().item():.3f})", flush=True)
    if err > 0.02 * ref.abs().max().item() + 1e-3:
        # localise: per 32x32 block error map of the first tile
        e = (out.float() - ref).abs()
        blk = e[:128, :min(n, 256)].reshape(4, 32, -1, 32).amax(dim=(1, 3))
        print("block error map (rows=32-row groups, cols=32-col groups):\n", blk.cpu().numpy().round(3), flush=True)
        print("out[0,:8]", out[0, :8].float().cpu().numpy(), "\nref[0,:8]", ref[0, :8].cpu().numpy(), flush=True)
        raise SystemExit(1)


@stage
def gemm_tiny():
    _gemm_case(128, 128, 64)


@stage
def gemm_k256():
    _gemm_case(128, 128, 256)


@stage
def gemm_n256():
    _gemm_case(128, 256, 512)


@stage
def gemm_multi_tile():
    _gemm_case(1024, 1024, 1024)


@stage
def gemm_tails():
    _gemm_case(200, 160, 72)


@stage
def gemm_big():
    _gemm_case(36864, 2048, 4096)


@stage
def pair_tiny():
    os.environ["TP_GEMM_MODE"] = "2"
    _gemm_case(256, 256, 64)


@stage
def pair_k1024():
    os.environ["TP_GEMM_MODE"] = "2"
    _gemm_case(256, 256, 1024)


@stage
def pair_multi():
    os.environ["TP_GEMM_MODE"] = "2"
    _gemm_case(2048, 1024, 1024)


@stage
def pair_tails():
    os.environ["TP_GEMM_MODE"] = "2"
    _gemm_case(1000, 512, 200)


@stage
def pair_big():
    os.environ["TP_GEMM_MODE"] = "2"
    _gemm_case(36864, 2048, 4096)


@stage
def gemm_shapes_bench():
    """Per-shape throughput of every GEMM of the N=64, s=2, H=4096 forward, one-CTA vs CTA-pair kernels."""
    import torch
    from tokenpacker_b200.kernels import gemm_bf16
    shapes = [("kv_proj.0  ", 36864, 2048, 4096, True), ("kv_proj.2  ", 36864, 1024, 1024, False), ("q-side 1024", 9216, 1024, 1024, False),
              ("mlp.0      ", 9216, 4096, 1024, True), ("mlp.2      ", 9216, 4096, 4096, False)]
    for name, m, n, k, gelu in shapes:
        a = torch.randn(m, k, device="cuda").bfloat16()
        b = (torch.randn(n, k, device="cuda") * 0.02).bfloat16()
        bias = torch.randn(n, device="cuda")
        c = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        res = []
        for mode in ("1", "2"):
            os.environ["TP_GEMM_MODE"] = mode
            for _ in range(3):
                gemm_bf16(a, b, bias=bias, gelu=gelu, out=c)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                gemm_bf16(a, b, bias=bias, gelu=gelu, out=c)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            res.append(f"mode{mode}: {ms * 1e3:8.1f} us {2.0 * m * n * k / ms / 1e9:7.1f} TF/s")
        # cuBLAS reference point for the same shape (torch.matmul, no epilogue)
        for _ in range(3):
            torch.matmul(a, b.t())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            torch.matmul(a, b.t())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{name} M={m} N={n} K={k}: " + " | ".join(res) + f" | cuBLAS(no epilogue): {ms * 1e3:8.1f} us {2.0 * m * n * k / ms / 1e9:7.1f} TF/s", flush=True)


@stage
def projector_small():
    import numpy as np
    import torch
    from oracle import tokenpacker_oracle as tpo
    from tokenpacker_b200 import TokenPackerB200
    for s in (2, 3, 4):
        hidden, n = 128, 2
        params = {k: tpo.round_bf16(v) for k, v in tpo.make_params(hidden, seed=100 + s).items()}
        m = TokenPackerB200(hidden_size=hidden, scale_factor=s)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
        m = m.to("cuda", torch.bfloat16).eval()
        x0, xm = tpo.make_inputs(n, seed=200 + s)
        x0, xm = tpo.round_bf16(x0), tpo.round_bf16(xm)
        with torch.no_grad():
            out = m((torch.from_numpy(x0).cuda().bfloat16(), torch.from_numpy(xm).cuda().bfloat16()))
        torch.cuda.synchronize()
        ref = tpo.tokenpacker_forward(params, x0, xm, s)
        o = out.float().cpu().numpy().astype(np.float64)
        rel = np.sqrt(((o - ref) ** 2).mean()) / np.sqrt((ref ** 2).mean())
        print(f"projector s={s} H={hidden}: rel-rms {rel:.3e} max-abs {np.abs(o - ref).max():.3e} (ref rms {np.sqrt((ref**2).mean()):.3f})", flush=True)


@stage
def projector_timing():
    import torch
    from tokenpacker_b200 import TokenPackerB200
    from oracle import tokenpacker_oracle as tpo
    m = TokenPackerB200(hidden_size=4096, scale_factor=2).to("cuda", torch.bfloat16).eval()
    n = 64
    x0 = torch.randn(n, 576, 1024, device="cuda").bfloat16()
    xm = torch.randn(n, 576, 4096, device="cuda").bfloat16()
    with torch.no_grad():
        for _ in range(3):
            m((x0, xm))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            m((x0, xm))
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"N=64 s=2 H=4096: {ms:.3f} ms/call -> {n * 144 / ms * 1e3:.3e} tok/s, {tpo.flops_per_crop(2) * n / ms / 1e9:.1f} TFLOP/s", flush=True)


@stage
def train_step_timing():
    import torch
    from tokenpacker_b200 import TokenPackerB200
    from tokenpacker_b200 import synthetic as syn
    m = TokenPackerB200(hidden_size=4096, scale_factor=2).to("cuda", torch.bfloat16).train()
    n = 64
    x0 = torch.randn(n, 576, 1024, device="cuda").bfloat16()
    xm = torch.randn(n, 576, 4096, device="cuda").bfloat16()
    gw = torch.randn(n, 144, 4096, device="cuda").bfloat16()
    def step():
        for p in m.parameters():
            p.grad = None
        out = m((x0, xm))
        out.backward(gw)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = syn.flops_per_crop(2) * n
    print(f"train step (fwd+bwd, N=64 s=2 H=4096): {ms:.3f} ms; fwd F_alg x3 = {3 * fl / 1e12:.2f} TF -> {3 * fl / ms / 1e9:.0f} TF/s equivalent; "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB", flush=True)
    # context: PyTorch eager autograd over the reference's op sequence (oracle/torch_port.py), bf16, same GPU
    from oracle import torch_port
    pd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    def ref_step():
        for p in pd.values():
            p.grad = None
        torch_port.forward(pd, x0, xm, 2).backward(gw)
    for _ in range(2):
        ref_step()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        ref_step()
    e1.record()
    torch.cuda.synchronize()
    print(f"reference op sequence, eager autograd bf16 on the same GPU: {e0.elapsed_time(e1) / 5:.3f} ms per fwd+bwd step", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        STAGES[sys.argv[1]]()
        sys.exit(0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "first_light.log"), "w")
    for name in STAGES:
        try:
            r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=300)
            msg = f"=== {name}: exit {r.returncode}\n{r.stdout}{r.stderr[-3000:] if r.returncode else ''}\n"
        except subprocess.TimeoutExpired as e:
            msg = f"=== {name}: TIMEOUT\n{e.stdout or ''}\n"
        print(msg, flush=True)
        log.write(msg)
        log.flush()
