"""Time the forward's GEMM shapes through tp_gemm_bf16 of a given build (A/B of epilogue variants):
    python tools/shape_times.py tokenpacker_b200/libtokenpacker_b200.so [more.so ...]"""
import ctypes as C
import os
import sys

import torch

shapes = [("kv_proj.0", 36864, 2048, 4096, 1), ("kv_proj.2", 36864, 1024, 1024, 0), ("q-side", 9216, 1024, 1024, 0),
          ("mlp.0", 9216, 4096, 1024, 1), ("mlp.2", 9216, 4096, 4096, 0)]
bufs = {}
for name, m, n, k, gelu in shapes:
    torch.manual_seed(0)
    bufs[name] = (torch.randn(m, k, device="cuda").bfloat16(), (torch.randn(n, k, device="cuda") * 0.02).bfloat16(),
                  torch.randn(n, device="cuda"), torch.empty(m, n, device="cuda", dtype=torch.bfloat16))
ref = {}
for path in sys.argv[1:]:
    lib = C.CDLL(os.path.abspath(path))
    lib.tp_gemm_bf16.restype = C.c_int
    lib.tp_gemm_bf16.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                 C.c_void_p, C.c_int, C.c_float, C.c_void_p]
    line = []
    for name, m, n, k, gelu in shapes:
        a, b, bias, c = bufs[name]
        s = torch.cuda.current_stream().cuda_stream
        call = lambda: lib.tp_gemm_bf16(a.data_ptr(), k, b.data_ptr(), k, c.data_ptr(), n, m, n, k, bias.data_ptr(), gelu, 1.0, s)
        for _ in range(5):
            assert call() == 0
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(30):
                call()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 30 * 1e3)
        same = ""
        if name in ref:
            same = " same-bits" if torch.equal(ref[name], c) else " DIFFERENT-BITS"
        else:
            ref[name] = c.clone()
        line.append(f"{name} {best:6.1f}us{same}")
    print(os.path.basename(path), " | ".join(line))
