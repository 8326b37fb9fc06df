"""Splice of the projected visual tokens into the LLM's input embeddings.

Mirrors ``LlavaMetaForCausalLM.prepare_inputs_labels_for_multimodal`` (llava/model/llava_arch.py:100-233) for the
configuration the released recipes use (``mm_use_im_start_end = False``): every IMAGE_TOKEN_INDEX placeholder of a sample is
replaced by the next image's visual rows, text tokens are looked up in the embedding table, sequences are right-padded to
the longest one, labels get IGNORE_INDEX over visual / padded positions, the attention mask follows the reference's rule.
The reference does this with Python lists and dozens of small ``torch.cat`` kernels per sample; here the host builds one index
vector (it has to look at the token ids anyway) and ONE gather kernel (tp_gather_rows) writes the [B, Lmax, H] buffer.
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


def splice_plan(input_ids, cu_seqlens, labels=None, attention_mask=None) -> SplicePlan:
    """Host-side plan.  input_ids [B, L]; cu_seqlens [n_images + 1]: row ranges of the image sequences inside the packed
    visual rows, consumed in order — one per image token, and one by every sample WITHOUT an image token (llava_arch.py:121-134)."""
    ids = np.asarray(input_ids, dtype=np.int64)
    cu = [int(v) for v in cu_seqlens]
    B, L = ids.shape
    lab = None if labels is None else np.asarray(labels, dtype=np.int64)
    rows, lrows, img = [], [], 0
    for b in range(B):
        cur, lcur = [], []
        pos = np.where(ids[b] == IMAGE_TOKEN_INDEX)[0]
        if pos.size == 0:
            cur.append(ids[b])
            if lab is not None:
                lcur.append(lab[b])
            img += 1
        else:
            start = 0
            for p in pos.tolist():
                if img + 1 >= len(cu):
                    raise ValueError("more image tokens than image sequences")
                cur.append(ids[b, start:p])
                cur.append(-(np.arange(cu[img], cu[img + 1], dtype=np.int64)) - 2)
                if lab is not None:
                    lcur.append(lab[b, start:p])
                    lcur.append(np.full(cu[img + 1] - cu[img], IGNORE_INDEX, dtype=np.int64))
                img += 1
                start = p + 1
            cur.append(ids[b, start:])
            if lab is not None:
                lcur.append(lab[b, start:])
        rows.append(np.concatenate(cur))
        if lab is not None:
            lrows.append(np.concatenate(lcur))
    lengths = [int(r.shape[0]) for r in rows]
    lmax = max(lengths)
    src = np.full((B, lmax), -1, dtype=np.int64)
    for b, r in enumerate(rows):
        src[b, :r.shape[0]] = r
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
    return SplicePlan(src.reshape(-1), lengths, lmax, out_labels, out_mask)


def splice_multimodal(input_ids: torch.Tensor, embed_weight: torch.Tensor, visual_rows: torch.Tensor, cu_seqlens, labels=None,
                      attention_mask=None):
    """Returns (attention_mask, inputs_embeds [B, Lmax, H], labels) like llava_arch.py:233 (its None / past_key_values slots dropped).

    embed_weight: the LLM's ``embed_tokens.weight`` [V, H]; visual_rows: packed visual tokens [sum L_i, H] (e.g. the output of
    ``TokenPackerB200.forward_packed``, or ``projector(x).flatten(0, 1)`` with cu_seqlens = arange * M); both CUDA."""
    if not (embed_weight.is_cuda and visual_rows.is_cuda):
        raise RuntimeError("tokenpacker_b200 has no CPU path: embed_weight and visual_rows must be CUDA tensors")
    device = visual_rows.device
    hidden = int(embed_weight.shape[1])
    if visual_rows.dim() != 2 or visual_rows.shape[1] != hidden:
        raise ValueError("visual_rows must be [rows, hidden]")
    plan = splice_plan(input_ids.cpu().numpy(), [int(v) for v in cu_seqlens],
                       None if labels is None else labels.cpu().numpy(), None if attention_mask is None else attention_mask.cpu().numpy())
    if plan.src_index.max(initial=-1) >= embed_weight.shape[0]:
        raise ValueError("token id outside the embedding table")
    B = int(input_ids.shape[0])
    table = embed_weight.detach().to(torch.bfloat16).contiguous()
    vis = visual_rows.detach().to(torch.bfloat16).contiguous()
    with torch.cuda.device(device):
        src = torch.from_numpy(plan.src_index).to(device)
        out = torch.empty((B, plan.lmax, hidden), dtype=torch.bfloat16, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        check(lib.tp_gather_rows(table.data_ptr(), vis.data_ptr(), hidden, src.data_ptr(), src.numel(), out.data_ptr(), stream), "tp_gather_rows")
    new_labels = None if plan.labels is None else torch.from_numpy(plan.labels).to(device=labels.device, dtype=labels.dtype)
    new_mask = None if plan.attention_mask is None else torch.from_numpy(plan.attention_mask).to(device=attention_mask.device)
    return new_mask, (out if embed_weight.dtype == torch.bfloat16 else out.to(embed_weight.dtype)), new_labels
