"""Where do the warp roles of the CTA-pair GEMM spend their cycles?  (instrumented build: make -C tokenpacker_b200/csrc prof)

    python tools/gemm_phase_profile.py        # on a B200; prints per-shape averages over CTAs, in cycles and % of role total
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

LIBNAME = sys.argv[1] if len(sys.argv) > 1 else "libtokenpacker_b200_prof.so"
print("==", LIBNAME)
lib = C.CDLL(os.path.join(ROOT, "tokenpacker_b200", LIBNAME))
lib.tp_gemm_bf16_prof.restype = C.c_int
lib.tp_gemm_bf16_prof.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                  C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
os.environ["TP_GEMM_MODE"] = "2"
shapes = [("kv_proj.0", 36864, 2048, 4096, 1), ("kv_proj.2", 36864, 1024, 1024, 0), ("q-side", 9216, 1024, 1024, 0),
          ("mlp.0", 9216, 4096, 1024, 1), ("mlp.2", 9216, 4096, 4096, 0)]
for name, m, n, k, gelu in shapes:
    a = torch.randn(m, k, device="cuda").bfloat16()
    b = (torch.randn(n, k, device="cuda") * 0.02).bfloat16()
    bias = torch.randn(n, device="cuda")
    c = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    prof = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        st = lib.tp_gemm_bf16_prof(a.data_ptr(), k, b.data_ptr(), k, c.data_ptr(), n, m, n, k, bias.data_ptr(), gelu, 1.0, prof.data_ptr(), s)
        assert st == 0, st
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lib.tp_gemm_bf16_prof(a.data_ptr(), k, b.data_ptr(), k, c.data_ptr(), n, m, n, k, bias.data_ptr(), gelu, 1.0, prof.data_ptr(), s)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"   timed: {us:.1f} us  {2.0 * m * n * k / us / 1e6:.1f} TF/s")
    p = prof.cpu().reshape(148, 16).double()
    lead, peer = p[0::2], p[1::2]
    tiles = ((m + 255) // 256) * (n // 256)
    ideal = tiles / 74 * (k / 64) * 512
    print(f"{name}: M={m} N={n} K={k} tiles/pair={tiles / 74:.2f} ideal MMA cycles/pair={ideal:.0f}")
    print(f"   producer(leader): wait-empty {lead[:, 0].mean():9.0f} of {lead[:, 1].mean():9.0f} ({100 * lead[:, 0].mean() / lead[:, 1].mean():.0f}%)"
          f" | producer(peer): wait-empty {peer[:, 0].mean():9.0f} of {peer[:, 1].mean():9.0f}")
    print(f"   mma: wait-full {lead[:, 2].mean():9.0f} ({100 * lead[:, 2].mean() / lead[:, 4].mean():.0f}%)  wait-tmem {lead[:, 3].mean():9.0f} "
          f"({100 * lead[:, 3].mean() / lead[:, 4].mean():.0f}%)  total {lead[:, 4].mean():9.0f}  -> issue+other {lead[:, 4].mean() - lead[:, 2].mean() - lead[:, 3].mean():9.0f}")
    print(f"   epilogue warp0: wait-acc {p[:, 5].mean():9.0f}  busy {p[:, 6].mean():9.0f}  busy/tile {p[:, 6].mean() / (tiles / 74):7.0f}"
          f"  of which per tile: tcgen05.wait::ld {p[:, 8].mean() / (tiles / 74):6.0f}  fence.proxy.async {p[:, 9].mean() / (tiles / 74):6.0f}"
          f"  wait_group.read+barrier {p[:, 10].mean() / (tiles / 74):6.0f}")
