"""Multi-rank host logic on CPU: world_size 2, gloo, 127.0.0.1 — sharding bounds, ragged all-gather, and that gathered
blocks reassemble (oracle restatement of llava_arch.py:139-155) to the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import hd_oracle as hdo


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_crops, m, hdim, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tokenpacker_b200.dist import all_gather_crops, shard_bounds, shard_counts
    try:
        full = torch.from_numpy(np.random.default_rng(0).standard_normal((n_crops, m, hdim)).astype(np.float32))
        lo, hi = shard_bounds(n_crops, world, rank)
        counts = shard_counts(n_crops, world)
        assert sum(counts) == n_crops and counts[rank] == hi - lo
        local = full[lo:hi] * 2.0 + 1.0                       # stand-in for the per-crop projection (crop-independent)
        gathered = all_gather_crops(local, counts)
        assert torch.equal(gathered, full * 2.0 + 1.0)
        np.save(os.path.join(tmp, f"g{rank}.npy"), gathered.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_crops", [7, 8])        # ragged (4+3) and even (4+4) shards
def test_two_rank_gather_and_assembly(tmp_path, n_crops):
    world, m, hdim = 2, 3, 8
    mp.spawn(_worker, args=(world, _free_port(), n_crops, m, hdim, str(tmp_path)), nprocs=world, join=True)
    g0, g1 = np.load(tmp_path / "g0.npy"), np.load(tmp_path / "g1.npy")
    np.testing.assert_array_equal(g0, g1)
    # images whose crops straddle the rank boundary still assemble correctly from the gathered blocks
    grids = [(2, 2), (1, 1), (1, 1)] if n_crops == 7 else [(2, 3), (1, 1)]
    assert sum(hdo.n_crops(a, b) for a, b in grids) == n_crops
    sep, ret = np.full(hdim, -1.0, np.float32), np.full(hdim, -2.0, np.float32)
    packed, cu = hdo.hd_assemble(g0, [a for a, _ in grids], [b for _, b in grids], sep, ret)
    full = np.random.default_rng(0).standard_normal((n_crops, m, hdim)).astype(np.float32) * 2.0 + 1.0
    ref, ref_cu = hdo.hd_assemble(full, [a for a, _ in grids], [b for _, b in grids], sep, ret)
    np.testing.assert_array_equal(packed, ref)
    np.testing.assert_array_equal(cu, ref_cu)


def test_shard_bounds_partition():
    from tokenpacker_b200.dist import shard_bounds
    for n in (1, 5, 8, 231, 256, 449):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
