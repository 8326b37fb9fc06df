"""Seconds-long back-to-back runs of one GEMM shape: cuBLAS (torch.matmul) vs ours, with clock/power sampling.
Answers: are we power-limited, and how does energy efficiency compare?"""
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from tokenpacker_b200.kernels import gemm_bf16  # noqa: E402

m, n, k = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (36864, 2048, 4096)
secs = float(sys.argv[4]) if len(sys.argv) > 4 else 3.0
a = torch.randn(m, k, device="cuda").bfloat16()
b = (torch.randn(n, k, device="cuda") * 0.02).bfloat16()
bias = torch.randn(n, device="cuda")
c = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)


def sample(stop, rows):
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu", "--format=csv,noheader,nounits", "-lms", "50"],
                         stdout=subprocess.PIPE, text=True)
    for line in p.stdout:
        rows.append(line.strip())
        if stop.is_set():
            break
    p.terminate()


def run(name, fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    stop, rows = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, rows), daemon=True)
    th.start()
    time.sleep(0.2)
    iters = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < secs:
        for _ in range(50):
            fn()
        iters += 50
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    stop.set()
    time.sleep(0.2)
    vals = [[float(x) for x in r.split(",")] for r in rows[4:] if r.count(",") == 2]
    clk = sorted(v[0] for v in vals)[len(vals) // 2] if vals else -1
    pw = sorted(v[1] for v in vals)[len(vals) // 2] if vals else -1
    tmp = max(v[2] for v in vals) if vals else -1
    print(f"{name:28s} {ms * 1e3:8.1f} us  {2.0 * m * n * k / ms / 1e9:7.1f} TF/s   median clk {clk:.0f} MHz  power {pw:.0f} W  temp {tmp:.0f}C  ({len(vals)} samples)", flush=True)
    time.sleep(1.0)


print(f"shape M={m} N={n} K={k}, {secs}s each")
run("cuBLAS (no epilogue)", lambda: torch.matmul(a, b.t(), out=c))
os.environ["TP_GEMM_MODE"] = "2"
run("ours pair, bias", lambda: gemm_bf16(a, b, bias=bias, out=c))
run("ours pair, bias+gelu", lambda: gemm_bf16(a, b, bias=bias, gelu=True, out=c))
os.environ["TP_GEMM_MODE"] = "1"
run("ours 1-CTA, bias", lambda: gemm_bf16(a, b, bias=bias, out=c))
run("cuBLAS again", lambda: torch.matmul(a, b.t(), out=c))
