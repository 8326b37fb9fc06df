"""2-rank NCCL run of the sharded HD path (only runs when >= 2 GPUs are visible): every rank projects its crop block,
all-gathers over NVLink and assembles; the result must equal the single-GPU forward_packed bit for bit."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    from tokenpacker_b200 import TokenPackerB200
    from tokenpacker_b200 import synthetic as syn
    from tokenpacker_b200.dist import FusedGatherTokenPacker, ShardedTokenPacker, shard_bounds, shard_counts
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        s, hidden = 4, 256
        grids = [(2, 2), (1, 1), (1, 3)]                      # 5 + 1 + 4 = 10 crops; image 0 straddles the rank boundary
        n = 10
        m = TokenPackerB200(hidden_size=hidden, scale_factor=s)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in syn.synthetic_state_dict(hidden, seed=3).items()})
        m = m.to(f"cuda:{rank}", torch.bfloat16).eval()
        g = torch.Generator(device=f"cuda:{rank}").manual_seed(7)       # same stream on every rank -> same global batch
        x0 = torch.randn(n, 576, 1024, device=f"cuda:{rank}", generator=g).bfloat16()
        xm = torch.randn(n, 576, 4096, device=f"cuda:{rank}", generator=g).bfloat16()
        sep = torch.randn(hidden, device=f"cuda:{rank}", generator=g).bfloat16()
        ret = torch.randn(hidden, device=f"cuda:{rank}", generator=g).bfloat16()
        hb, wb = [a for a, _ in grids], [b for _, b in grids]
        lo, hi = shard_bounds(n, world, rank)
        with torch.no_grad():
            packed, cu = ShardedTokenPacker(m).forward_hd((x0[lo:hi], xm[lo:hi]), shard_counts(n, world), hb, wb, sep, ret)
            ref, ref_cu = m.forward_packed((x0, xm), hb, wb, sep, ret)
        assert torch.equal(cu, ref_cu)
        assert torch.equal(packed, ref), "sharded + all-gather + assembly differs from the single-GPU packed forward"
        # fused path: the last GEMM's TMA stores write into every rank's gathered buffer over NVLink (no NCCL all-gather)
        fused = FusedGatherTokenPacker(m)
        with torch.no_grad():
            for _ in range(3):                    # repeated use of the same symmetric buffer
                packed_f, cu_f = fused.forward_hd((x0[lo:hi], xm[lo:hi]), shard_counts(n, world), hb, wb, sep, ret)
                assert torch.equal(cu_f, ref_cu)
                assert torch.equal(packed_f, ref), "fused peer-store all-gather differs from the single-GPU packed forward"
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_gpu_hd_allgather(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
