"""Single-GPU probe of the fused exchange's store cost: the configs[4] per-rank share (32 crops, s=4, H=4096) written into the packed rows
of 1 / 2 / 4 / 8 'peer' buffers that all live on THIS GPU.  No NVLink involved: the time added per extra destination is the cost of
issuing and draining the extra TMA stores (pieces of the 37-row crop stride x destinations), not of the links."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tokenpacker_b200 import TokenPackerB200

def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    s, hidden, n_local, n_total = 4, 4096, 32, 256
    m = TokenPackerB200(scale_factor=s, hidden_size=hidden).to(dev, torch.bfloat16).eval()
    x0 = torch.randn(n_local, 576, 1024, device=dev).to(torch.bfloat16)
    xm = torch.randn(n_local, 576, 4096, device=dev).to(torch.bfloat16)
    mq = m.num_queries
    bufs = [torch.zeros(n_total * (mq + 1), hidden, dtype=torch.bfloat16, device=dev) for _ in range(8)]
    res = {}
    with torch.no_grad():
        for n_dst in (1, 2, 4, 8, 1, 8):
            ptrs = [b.data_ptr() for b in bufs[:n_dst]]
            for _ in range(5):
                m.forward_into_peers((x0, xm), ptrs, 64, out_crop_rows=mq + 1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(40):
                m.forward_into_peers((x0, xm), ptrs, 64, out_crop_rows=mq + 1)
            e1.record()
            torch.cuda.synchronize()
            res.setdefault(n_dst, []).append(round(e0.elapsed_time(e1) / 40, 4))
        same = all(torch.equal(bufs[0], b) for b in bufs[1:])
        dense = m((x0, xm))                                   # [32, M, H]: the packed rows must hold exactly these bits
        rows = bufs[0].view(n_total, mq + 1, hidden)[64:64 + n_local, :mq]
        same = same and torch.equal(rows, dense)
    print(json.dumps({"workload": "32 crops, s=4, H=4096 into packed rows (crop stride 37) of n local destinations", "ms_by_destinations": res,
                      "all_destinations_equal_and_match_dense_forward": same, "TP_SCHEDULE": os.environ.get("TP_SCHEDULE", "0")}))

if __name__ == "__main__":
    main()
