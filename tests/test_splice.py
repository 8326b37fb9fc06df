"""Text/vision splice: oracle and product plan vs fixtures produced by the reference's own prepare_inputs_labels_for_multimodal
(llava_arch.py:100-233, oracle/gen_golden.py:gen_splice); the gather kernel itself is checked on the GPU."""
import os

import numpy as np
import pytest

from oracle import hd_oracle as hdo
from oracle import splice_oracle as spo

CASES = ["equal", "ragged", "infer", "slice", "startend", "startend_ragged"]      # startend*: llava_arch.py:162-170 branch


def _image_seqs(g, name):
    feats = g[f"{name}_feats"]
    if f"{name}_grids" in g.files:
        grids = g[f"{name}_grids"].tolist()
        sep, ret = g["table"][int(g["sep_id"])], g["table"][int(g["ret_id"])]
        packed, cu = hdo.hd_assemble(feats, [a for a, _ in grids], [b for _, b in grids], sep, ret)
        return [packed[cu[i]:cu[i + 1]] for i in range(len(grids))]
    return [feats[i] for i in range(feats.shape[0])]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "splice.npz"))
    ids = g[f"{name}_ids"]
    labels = ids.copy() if f"{name}_labels" in g.files else None
    mask, embeds, new_labels = spo.splice(ids, np.ones_like(ids, dtype=bool), labels, _image_seqs(g, name), g["table"],
                                          im_start_end=name.startswith("startend"))
    np.testing.assert_array_equal(embeds, g[f"{name}_embeds"])
    np.testing.assert_array_equal(mask, g[f"{name}_mask"])
    if labels is not None:
        np.testing.assert_array_equal(new_labels, g[f"{name}_labels"])


@pytest.mark.parametrize("name", CASES)
def test_product_plan_matches_reference(golden_dir, name):
    """The product's host planner (pure index arithmetic, no CUDA) applied with numpy reproduces the reference outputs."""
    from tokenpacker_b200 import splice_plan
    g = np.load(os.path.join(golden_dir, "splice.npz"))
    ids = g[f"{name}_ids"]
    seqs = _image_seqs(g, name)
    visual = np.concatenate(seqs, axis=0)
    cu = np.concatenate([[0], np.cumsum([s.shape[0] for s in seqs])])
    labels = ids.copy() if f"{name}_labels" in g.files else None
    plan = splice_plan(ids, cu, labels, np.ones_like(ids, dtype=bool), im_start_end=name.startswith("startend"))
    src = plan.src_index
    rows = np.zeros((src.shape[0], g["table"].shape[1]), dtype=np.float32)
    rows[src >= 0] = g["table"][src[src >= 0]]
    rows[src <= -2] = visual[-src[src <= -2] - 2]
    np.testing.assert_array_equal(rows.reshape(ids.shape[0], plan.lmax, -1), g[f"{name}_embeds"])
    np.testing.assert_array_equal(plan.attention_mask, g[f"{name}_mask"])
    if labels is not None:
        np.testing.assert_array_equal(plan.labels, g[f"{name}_labels"])


def test_ragged_without_labels_is_rejected_like_the_reference():
    from tokenpacker_b200 import splice_plan
    with pytest.raises(ValueError):
        splice_plan(np.array([[1, -200, 3], [4, 5, 6]]), [0, 4, 8], None, np.ones((2, 3), dtype=bool))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gather_kernel_matches_reference(golden_dir, name):
    import torch
    from tokenpacker_b200 import splice_multimodal
    g = np.load(os.path.join(golden_dir, "splice.npz"))
    ids = torch.from_numpy(g[f"{name}_ids"])
    seqs = _image_seqs(g, name)
    visual = torch.from_numpy(np.concatenate(seqs, axis=0)).cuda().bfloat16()
    cu = np.concatenate([[0], np.cumsum([s.shape[0] for s in seqs])])
    table = torch.from_numpy(g["table"]).cuda().bfloat16()
    labels = ids.clone() if f"{name}_labels" in g.files else None
    mask, embeds, new_labels = splice_multimodal(ids.cuda(), table, visual, cu, None if labels is None else labels.cuda(),
                                                 torch.ones_like(ids, dtype=torch.bool).cuda(), im_start_end=name.startswith("startend"))
    ref = torch.from_numpy(g[f"{name}_embeds"]).bfloat16()          # pure data movement: bit-exact on the bf16-rounded values
    assert torch.equal(embeds.cpu(), ref)
    np.testing.assert_array_equal(mask.cpu().numpy(), g[f"{name}_mask"])
    if labels is not None:
        np.testing.assert_array_equal(new_labels.cpu().numpy(), g[f"{name}_labels"])


def test_start_end_branch_only_differs_in_labels(golden_dir):
    """llava_arch.py:162-170 places the same rows as :171-179; after the visual rows it keeps the label of the placeholder
    position (not of <im_end>)."""
    from tokenpacker_b200 import splice_plan
    ids = np.array([[1, 30, -200, 31, 4, 7]])
    a = splice_plan(ids, [0, 3], ids.copy(), None, im_start_end=False)
    b = splice_plan(ids, [0, 3], ids.copy(), None, im_start_end=True)
    np.testing.assert_array_equal(a.src_index, b.src_index)
    np.testing.assert_array_equal(a.labels, [[1, 30, -100, -100, -100, 31, 4, 7]])
    np.testing.assert_array_equal(b.labels, [[1, 30, -100, -100, -100, -200, 4, 7]])
    np.testing.assert_array_equal(b.table_grad, [False, True, False, False, False, True, False, False])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ragged", "slice", "startend_ragged"])
def test_splice_backward_matches_autograd(golden_dir, name):
    """Gradients w.r.t. the visual rows (the projector's output) and the embedding table vs torch autograd through an
    index-based restatement; in the start/end branch only the <im_start>/<im_end> table rows receive gradient (.detach() at
    llava_arch.py:163,184)."""
    import torch
    from tokenpacker_b200 import splice_multimodal, splice_plan
    g = np.load(os.path.join(golden_dir, "splice.npz"))
    start_end = name.startswith("startend")
    ids = torch.from_numpy(g[f"{name}_ids"])
    seqs = _image_seqs(g, name)
    cu = np.concatenate([[0], np.cumsum([s.shape[0] for s in seqs])])
    visual = torch.from_numpy(np.concatenate(seqs, axis=0)).cuda().bfloat16().requires_grad_(True)
    table = torch.from_numpy(g["table"]).cuda().bfloat16().requires_grad_(True)
    _, embeds, _ = splice_multimodal(ids.cuda(), table, visual, cu, ids.clone().cuda(), None, im_start_end=start_end)
    torch.manual_seed(3)
    w = torch.randn(embeds.shape, device="cuda").bfloat16()
    (embeds.float() * w.float()).sum().backward()
    g_vis, g_tab = visual.grad.clone(), table.grad.clone()

    plan = splice_plan(ids.numpy(), cu, None, None, im_start_end=start_end)
    src = torch.from_numpy(plan.src_index).cuda()
    v2 = visual.detach().clone().requires_grad_(True)
    t2 = table.detach().clone().requires_grad_(True)
    hidden = t2.shape[1]
    rows = torch.zeros((src.numel(), hidden), device="cuda")
    text = src >= 0
    live = text if plan.table_grad is None else text & torch.from_numpy(plan.table_grad).cuda()
    dead = text & ~live
    rows = rows.index_put((torch.nonzero(live)[:, 0],), t2.float()[src[live]])
    rows = rows.index_put((torch.nonzero(dead)[:, 0],), t2.detach().float()[src[dead]])
    rows = rows.index_put((torch.nonzero(src <= -2)[:, 0],), v2.float()[-src[src <= -2] - 2])
    assert torch.equal(rows.view(embeds.shape).bfloat16(), embeds.detach())
    (rows.view(embeds.shape) * w.float()).sum().backward()
    assert torch.equal(g_vis, v2.grad.bfloat16())
    torch.testing.assert_close(g_tab.float(), t2.grad.float(), rtol=1e-2, atol=1e-2)
    if start_end:
        # only <im_start>/<im_end> rows (30, 31) — plus the ids of image-free samples, which llava_arch.py:121-134 embeds without
        # .detach() — receive gradient
        touched = {30, 31}
        for row in ids.tolist():
            if -200 not in row:
                touched.update(row)
        untouched = torch.ones(table.shape[0], dtype=torch.bool)
        untouched[sorted(touched)] = False
        assert torch.count_nonzero(g_tab[untouched.cuda()]) == 0
        assert torch.count_nonzero(g_tab[30]) > 0 and torch.count_nonzero(g_tab[31]) > 0
