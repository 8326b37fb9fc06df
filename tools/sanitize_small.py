"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from tokenpacker_b200 import TokenPackerB200, hd_tile, hd_tile_batch  # noqa: E402
from tokenpacker_b200.kernels import gemm_bf16, gemm_tn_bf16  # noqa: E402

torch.manual_seed(0)
a = torch.randn(300, 200, device="cuda").bfloat16()
b = torch.randn(512, 200, device="cuda").bfloat16()
gemm_bf16(a, b, bias=torch.randn(512, device="cuda"), gelu=True)
gemm_tn_bf16(torch.randn(200, 304, device="cuda").bfloat16(), torch.randn(200, 256, device="cuda").bfloat16())
for s, hidden in ((2, 256), (3, 128), (6, 128)):       # 6: the streamed-window attention kernels (forward + backward)
    m = TokenPackerB200(hidden_size=hidden, scale_factor=s).to("cuda", torch.bfloat16)
    x0 = torch.randn(2, 577, 1024, device="cuda").bfloat16()[:, 1:]
    xm = torch.randn(2, 577, 4096, device="cuda").bfloat16()[:, 1:]
    with torch.no_grad():
        out = m((x0, xm))
        packed, cu = m.forward_packed((x0, xm), [1, 1], [1, 1], torch.randn(hidden, device="cuda"), torch.randn(hidden, device="cuda"))
    tr = m((x0, xm))
    tr.float().pow(2).mean().backward()
os.environ["TP_GEMM_MODE"] = "2"            # CTA-pair kernel: the packed-row (3-D clipped box) TMA stores
m = TokenPackerB200(hidden_size=256, scale_factor=4).to("cuda", torch.bfloat16)
with torch.no_grad():
    packed, cu = m.forward_packed((torch.randn(4, 576, 1024, device="cuda").bfloat16(), torch.randn(4, 576, 4096, device="cuda").bfloat16()),
                                  [1, 1], [2, 1], torch.randn(256, device="cuda"), torch.randn(256, device="cuda"))
os.environ.pop("TP_GEMM_MODE")
crops, hb, wb = hd_tile(torch.randn(1, 3, 500, 700, device="cuda"), 9)
bcrops, _, _ = hd_tile_batch([torch.randn(3, 500, 700, device="cuda"), torch.randn(3, 300, 200, device="cuda")], 9)
torch.cuda.synchronize()
print("sanitize run complete", out.shape, packed.shape, crops.shape)
