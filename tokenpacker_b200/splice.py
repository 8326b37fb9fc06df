"""Splice of the projected visual tokens into the LLM's input embeddings.

Mirrors ``LlavaMetaForCausalLM.prepare_inputs_labels_for_multimodal`` (llava/model/llava_arch.py:100-233): every
IMAGE_TOKEN_INDEX placeholder of a sample is replaced by the next image's visual rows, text tokens are looked up in the
embedding table, sequences are right-padded to the longest one, labels get IGNORE_INDEX over visual / padded positions, the
attention mask follows the reference's rule.  ``im_start_end=True`` is the branch the reference takes when
``tune_mm_mlp_adapter and mm_use_im_start_end`` (:162-170): same embeddings, different label bookkeeping, and only the
<im_start>/<im_end> rows of the table receive gradient.  The reference does all this with Python lists and dozens of small
``torch.cat`` kernels per sample; here the host builds one index vector (it has to look at the token ids anyway) and ONE gather
kernel (tp_gather_rows) writes the [B, Lmax, H] buffer; the backward is the same kernel run with the inverse index.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from ._lib import lib, check

IGNORE_INDEX = -100          # llava/constants.py
IMAGE_TOKEN_INDEX = -200


@dataclass
class SplicePlan:
    src_index: np.ndarray            # int64 [B * Lmax]: >= 0 table row, -1 zero row, <= -2 visual row (-i - 2)
    lengths: list                    # spliced length of every sample
    lmax: int
    labels: np.ndarray | None        # int64 [B, Lmax]
    attention_mask: np.ndarray | None
    table_grad: np.ndarray | None = None    # bool [B * Lmax]: text positions whose table row receives gradient (None = all)


def splice_plan(input_ids, cu_seqlens, labels=None, attention_mask=None, im_start_end=False) -> SplicePlan:
    """Host-side plan.  input_ids [B, L]; cu_seqlens [n_images + 1]: row ranges of the image sequences inside the packed
    visual rows, consumed in order — one per image token, and one by every sample WITHOUT an image token (llava_arch.py:121-134)."""
    ids = np.asarray(input_ids, dtype=np.int64)
    cu = [int(v) for v in cu_seqlens]
    B, L = ids.shape
    lab = None if labels is None else np.asarray(labels, dtype=np.int64)
    rows, lrows, grows, img = [], [], [], 0
    for b in range(B):
        cur, lcur, gcur = [], [], []
        pos = np.where(ids[b] == IMAGE_TOKEN_INDEX)[0]
        if pos.size == 0:
            cur.append(ids[b])
            gcur.append(np.ones(L, dtype=bool))
            if lab is not None:
                lcur.append(lab[b])
            img += 1
        elif im_start_end:
            # llava_arch.py:162-170,176-177,183-184 restated slice for slice (python slice semantics, so a placeholder at
            # position 0 behaves as it does upstream); text outside <im_start>/<im_end> is detached there
            rest, lrest = ids[b], (lab[b] if lab is not None else None)
            while True:
                where = np.where(rest == IMAGE_TOKEN_INDEX)[0]
                if where.size == 0:
                    break
                p = int(where[0])
                if img + 1 >= len(cu):
                    raise ValueError("more image tokens than image sequences")
                n_vis = cu[img + 1] - cu[img]
                pieces = [(rest[:p - 1], False), (rest[p - 1:p], True),
                          (-(np.arange(cu[img], cu[img + 1], dtype=np.int64)) - 2, False), (rest[p + 1:p + 2], True)]
                for piece, grad in pieces:
                    cur.append(piece)
                    gcur.append(np.full(piece.shape[0], grad, dtype=bool))
                if lab is not None:
                    lcur += [lrest[:p], np.full(n_vis, IGNORE_INDEX, dtype=np.int64), lrest[p:p + 1]]
                    lrest = lrest[p + 2:]
                rest = rest[p + 2:]
                img += 1
            cur.append(rest)
            gcur.append(np.zeros(rest.shape[0], dtype=bool))
            if lab is not None:
                lcur.append(lrest)
        else:
            start = 0
            for p in pos.tolist():
                if img + 1 >= len(cu):
                    raise ValueError("more image tokens than image sequences")
                cur.append(ids[b, start:p])
                cur.append(-(np.arange(cu[img], cu[img + 1], dtype=np.int64)) - 2)
                gcur += [np.ones(p - start, dtype=bool), np.zeros(cu[img + 1] - cu[img], dtype=bool)]
                if lab is not None:
                    lcur.append(lab[b, start:p])
                    lcur.append(np.full(cu[img + 1] - cu[img], IGNORE_INDEX, dtype=np.int64))
                img += 1
                start = p + 1
            cur.append(ids[b, start:])
            gcur.append(np.ones(L - start, dtype=bool))
            if lab is not None:
                lcur.append(lab[b, start:])
        rows.append(np.concatenate(cur))
        grows.append(np.concatenate(gcur))
        if lab is not None:
            lrows.append(np.concatenate(lcur))
    lengths = [int(r.shape[0]) for r in rows]
    lmax = max(lengths)
    src = np.full((B, lmax), -1, dtype=np.int64)
    tgrad = np.zeros((B, lmax), dtype=bool)
    for b, r in enumerate(rows):
        src[b, :r.shape[0]] = r
        tgrad[b, :r.shape[0]] = grows[b]
    out_labels = None
    if lab is not None:
        out_labels = np.full((B, lmax), IGNORE_INDEX, dtype=np.int64)
        for b, r in enumerate(lrows):
            out_labels[b, :r.shape[0]] = r
    out_mask = None
    if attention_mask is not None:
        am = np.asarray(attention_mask)
        if len(set(lengths)) > 1:
            if lab is None:
                raise ValueError("ragged batch without labels: the reference itself cannot build the mask here (llava_arch.py:211-218)")
            out_mask = np.zeros((B, lmax), dtype=am.dtype)
            for b in range(B):                               # True x added tokens | old mask | False x right pad   (:211-219)
                added = lengths[b] - L
                out_mask[b, :added] = True
                out_mask[b, added:lengths[b]] = am[b]
        else:
            out_mask = np.concatenate([np.ones((B, lmax - L), dtype=am.dtype), am], axis=1)      # :226-229
    return SplicePlan(src.reshape(-1), lengths, lmax, out_labels, out_mask, tgrad.reshape(-1) if im_start_end else None)


def _gather(table, vis, hidden, src, out):
    stream = torch.cuda.current_stream(out.device).cuda_stream
    check(lib.tp_gather_rows(table.data_ptr(), vis.data_ptr(), hidden, src.data_ptr(), src.numel(), out.data_ptr(), stream), "tp_gather_rows")


_TABLE_CACHE: dict = {}


def _bf16_table(embed_weight: torch.Tensor) -> torch.Tensor:
    """bf16 view of the embedding table for the inference splice.  A bf16 table (every released recipe) is used in place; an fp16 /
    fp32 table would otherwise be re-cast on every call (0.5 GB for a 32k x 4096 fp32 table), so its bf16 copy is cached, keyed like
    the packed projector weights on (data_ptr, _version, dtype, shape).  Writes through a ``.data`` alias are not seen: inference only
    (the training path goes through _SpliceFunction, which casts per call)."""
    w = embed_weight.detach()
    if w.dtype == torch.bfloat16:
        return w.contiguous()
    key = (w.data_ptr(), embed_weight._version, w.dtype, tuple(w.shape), str(w.device))
    hit = _TABLE_CACHE.get("t")
    if hit is None or hit[0] != key:
        hit = (key, w.to(torch.bfloat16).contiguous())
        _TABLE_CACHE["t"] = hit
    return hit[1]


class _SpliceFunction(torch.autograd.Function):
    """out[i] = table[src[i]] | visual[-src[i]-2] | 0.  Every visual row is placed at most once, so its gradient is again a
    gather (inverse index, tp_gather_rows); table rows can repeat, so their gradient is an index_add over the text positions."""

    @staticmethod
    def forward(ctx, table, vis, src, inv_src, text_pos, text_ids, shape):
        hidden = table.shape[1]
        tb = table.detach().to(torch.bfloat16).contiguous()
        vb = vis.detach().to(torch.bfloat16).contiguous()
        out = torch.empty(shape, dtype=torch.bfloat16, device=vis.device)
        with torch.cuda.device(vis.device):
            _gather(tb, vb, hidden, src, out)
        ctx.save_for_backward(inv_src, text_pos, text_ids)
        ctx.meta = (table.shape, table.dtype, vis.shape, vis.dtype)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        inv_src, text_pos, text_ids = ctx.saved_tensors
        t_shape, t_dtype, v_shape, v_dtype = ctx.meta
        hidden = t_shape[1]
        g = grad_out.to(torch.bfloat16).contiguous().view(-1, hidden)
        g_table = g_vis = None
        if ctx.needs_input_grad[1]:
            gv = torch.empty(v_shape, dtype=torch.bfloat16, device=g.device)
            with torch.cuda.device(g.device):
                _gather(g, g, hidden, inv_src, gv)              # inv_src >= 0: row of grad_out; -1: visual row never placed
            g_vis = gv.to(v_dtype)
        if ctx.needs_input_grad[0]:
            g_table = torch.zeros(t_shape, dtype=torch.float32, device=g.device)
            g_table.index_add_(0, text_ids, g[text_pos].float())
            g_table = g_table.to(t_dtype)
        return g_table, g_vis, None, None, None, None, None


def splice_multimodal(input_ids: torch.Tensor, embed_weight: torch.Tensor, visual_rows: torch.Tensor, cu_seqlens, labels=None,
                      attention_mask=None, im_start_end=False):
    """Returns (attention_mask, inputs_embeds [B, Lmax, H], labels) like llava_arch.py:233 (its None / past_key_values slots dropped).

    embed_weight: the LLM's ``embed_tokens.weight`` [V, H]; visual_rows: packed visual tokens [sum L_i, H] (e.g. the output of
    ``TokenPackerB200.forward_packed``, or ``projector(x).flatten(0, 1)`` with cu_seqlens = arange * M); both CUDA.  Differentiable
    w.r.t. visual_rows (the projector's output) and embed_weight; ``im_start_end=True`` is the reference's
    ``tune_mm_mlp_adapter and mm_use_im_start_end`` branch (llava_arch.py:162-170)."""
    if not (embed_weight.is_cuda and visual_rows.is_cuda):
        raise RuntimeError("tokenpacker_b200 has no CPU path: embed_weight and visual_rows must be CUDA tensors")
    device = visual_rows.device
    hidden = int(embed_weight.shape[1])
    if visual_rows.dim() != 2 or visual_rows.shape[1] != hidden:
        raise ValueError("visual_rows must be [rows, hidden]")
    plan = splice_plan(input_ids.cpu().numpy(), [int(v) for v in cu_seqlens],
                       None if labels is None else labels.cpu().numpy(), None if attention_mask is None else attention_mask.cpu().numpy(),
                       im_start_end)
    if plan.src_index.max(initial=-1) >= embed_weight.shape[0]:
        raise ValueError("token id outside the embedding table")
    B = int(input_ids.shape[0])
    src_np = plan.src_index
    src = torch.from_numpy(src_np).to(device)
    needs_grad = torch.is_grad_enabled() and (embed_weight.requires_grad or visual_rows.requires_grad)
    if needs_grad:
        vis_pos = np.nonzero(src_np <= -2)[0]
        inv = np.full(visual_rows.shape[0], -1, dtype=np.int64)
        inv[-src_np[vis_pos] - 2] = vis_pos
        text = src_np >= 0 if plan.table_grad is None else (src_np >= 0) & plan.table_grad
        text_pos = np.nonzero(text)[0]
        out = _SpliceFunction.apply(embed_weight, visual_rows, src, torch.from_numpy(inv).to(device), torch.from_numpy(text_pos).to(device),
                                    torch.from_numpy(src_np[text_pos]).to(device), (B, plan.lmax, hidden))
    else:
        table = _bf16_table(embed_weight)
        vis = visual_rows.detach().to(torch.bfloat16).contiguous()
        out = torch.empty((B, plan.lmax, hidden), dtype=torch.bfloat16, device=device)
        with torch.cuda.device(device):
            _gather(table, vis, hidden, src, out)
    new_labels = None if plan.labels is None else torch.from_numpy(plan.labels).to(device=labels.device, dtype=labels.dtype)
    new_mask = None if plan.attention_mask is None else torch.from_numpy(plan.attention_mask).to(device=attention_mask.device)
    return new_mask, (out if embed_weight.dtype == torch.bfloat16 else out.to(embed_weight.dtype)), new_labels
