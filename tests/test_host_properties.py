"""Property tests (hypothesis) of the host-side index arithmetic: crop sharding, the HD assembly plan (tp_hd_plan) and the splice
plan.  Pure CPU; sizes are small so the whole file runs in seconds."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import hd_oracle as hdo


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 2000), world=st.integers(1, 16))
def test_shard_bounds_partition(n, world):
    from tokenpacker_b200.dist import shard_bounds, shard_counts
    bounds = [shard_bounds(n, world, r) for r in range(world)]
    assert bounds[0][0] == 0 and bounds[-1][1] == n
    for (lo, hi), (lo2, _) in zip(bounds, bounds[1:]):
        assert lo <= hi == lo2                                   # contiguous blocks, in rank order
    sizes = [hi - lo for lo, hi in bounds]
    assert sizes == shard_counts(n, world) and max(sizes) - min(sizes) <= 1 and sum(sizes) == n


grids = st.lists(st.tuples(st.integers(1, 5), st.integers(1, 5)), min_size=1, max_size=6)


@settings(max_examples=150, deadline=None)
@given(grids=grids, m=st.integers(1, 7))
def test_hd_plan_is_an_exact_partition_of_the_packed_rows(grids, m):
    """Crop segments, ',' rows and '\\n' rows tile [0, total) exactly once, per-image lengths follow llava_arch.py:139-155."""
    from tokenpacker_b200 import hd_plan, hd_seq_len
    hb, wb = [g[0] for g in grids], [g[1] for g in grids]
    plan = hd_plan(hb, wb, m)
    total = int(plan.cu_seqlens[-1])
    cover = np.zeros(total, dtype=np.int64)
    for r0 in plan.seg_row_offset.tolist():
        cover[r0:r0 + m] += 1
    cover[plan.sep_rows.numpy()] += 1
    cover[plan.ret_rows.numpy()] += 1
    assert (cover == 1).all()
    assert plan.n_crops == sum(hdo.n_crops(a, b) for a, b in grids) == plan.seg_row_offset.numel()
    lens = np.diff(plan.cu_seqlens.numpy()).tolist()
    assert lens == [hd_seq_len(a, b, m) for a, b in grids] == [hdo.hd_seq_len(a, b, m) for a, b in grids]
    assert plan.sep_rows.numel() == sum(a * (b - 1) for a, b in grids)
    assert plan.ret_rows.numel() == sum(a + (1 if a * b > 1 else 0) for a, b in grids)


@settings(max_examples=150, deadline=None)
@given(data=st.data(), start_end=st.booleans())
def test_splice_plan_invariants(data, start_end):
    from tokenpacker_b200 import splice_plan
    from tokenpacker_b200.splice import IGNORE_INDEX, IMAGE_TOKEN_INDEX
    B = data.draw(st.integers(1, 4))
    L = data.draw(st.integers(4, 12))
    ids = np.array(data.draw(st.lists(st.lists(st.integers(7, 50), min_size=L, max_size=L), min_size=B, max_size=B)))
    n_seq = 0
    for b in range(B):
        k = data.draw(st.integers(0, 2))
        pos = sorted(data.draw(st.lists(st.integers(1, L - 2), min_size=k, max_size=k, unique=True)))
        if start_end:                                   # keep placeholders apart: <im_start> IMG <im_end> triples do not overlap
            pos = [p for i, p in enumerate(pos) if i == 0 or p - pos[i - 1] >= 3]
        ids[b, pos] = IMAGE_TOKEN_INDEX
        n_seq += max(len(pos), 1)
    seq_lens = data.draw(st.lists(st.integers(1, 5), min_size=n_seq, max_size=n_seq))
    cu = np.concatenate([[0], np.cumsum(seq_lens)])
    labels = ids.copy()
    plan = splice_plan(ids, cu, labels, np.ones_like(ids, dtype=bool), im_start_end=start_end)
    src = plan.src_index.reshape(B, plan.lmax)
    vis = -src[src <= -2] - 2
    assert len(set(vis.tolist())) == vis.size                       # every visual row is placed at most once
    assert plan.lmax == max(plan.lengths)
    for b in range(B):
        row = src[b]
        assert (row[plan.lengths[b]:] == -1).all() and (row[:plan.lengths[b]] != -1).all()      # right padding only
        assert (plan.labels[b][row <= -1] == IGNORE_INDEX).all()    # visual and padded positions never carry a label
        n_text = int((ids[b] != IMAGE_TOKEN_INDEX).sum())
        assert int((row >= 0).sum()) == n_text                       # every text token embedded exactly once, in order
        np.testing.assert_array_equal(row[row >= 0], ids[b][ids[b] != IMAGE_TOKEN_INDEX])
    assert plan.attention_mask.shape == (B, plan.lmax)
