"""CPU oracle for the text/vision splice  --  TEST INFRASTRUCTURE ONLY (see tokenpacker_oracle.py).

numpy restatement of ``LlavaMetaForCausalLM.prepare_inputs_labels_for_multimodal`` (llava/model/llava_arch.py:100-233): each
IMAGE_TOKEN_INDEX placeholder of a sample is replaced by the next image's visual rows, text tokens are embedded by table
lookup, sequences are right-padded to the longest one.  ``im_start_end=True`` selects the branch taken when
``tune_mm_mlp_adapter and mm_use_im_start_end`` (:162-170,176-177), whose label bookkeeping differs (the reference's slices are
restated literally, including what they do when the placeholder is the first token).  Pinned by tests/golden/splice.npz,
produced by calling the reference method itself on a stand-in model (oracle/gen_golden.py).
"""
from __future__ import annotations

import numpy as np

IGNORE_INDEX = -100        # llava/constants.py
IMAGE_TOKEN_INDEX = -200


def splice(input_ids, attention_mask, labels, image_seqs, embed_table, im_start_end=False):
    """input_ids [B,L] int; attention_mask [B,L] bool or None; labels [B,L] int or None; image_seqs: list of [L_i,H] arrays,
    consumed in order (one per image token; a sample WITHOUT an image token still consumes one, llava_arch.py:121-134);
    embed_table [V,H].  Returns (attention_mask, inputs_embeds [B,Lmax,H], labels) like llava_arch.py:233."""
    input_ids = np.asarray(input_ids)
    B, L = input_ids.shape
    hdim = embed_table.shape[1]
    embeds, new_labels, img = [], [], 0
    for b in range(B):
        ids = input_ids[b]
        parts, lparts = [], []
        if (ids == IMAGE_TOKEN_INDEX).sum() == 0:                       # :121-134
            parts.append(embed_table[ids])
            if labels is not None:
                lparts.append(labels[b])
            img += 1
        else:
            cur_labels = labels[b] if labels is not None else None
            while True:                                                  # :139-181
                pos = np.where(ids == IMAGE_TOKEN_INDEX)[0]
                if pos.size == 0:
                    break
                p = int(pos[0])
                feat = image_seqs[img]
                img += 1
                if im_start_end:                                         # :162-170,176-177
                    parts.append(embed_table[ids[:p - 1]])
                    parts.append(embed_table[ids[p - 1:p]])
                    parts.append(feat)
                    parts.append(embed_table[ids[p + 1:p + 2]])
                    if labels is not None:
                        lparts.append(cur_labels[:p])
                        lparts.append(np.full(feat.shape[0], IGNORE_INDEX, dtype=labels.dtype))
                        lparts.append(cur_labels[p:p + 1])
                        cur_labels = cur_labels[p + 2:]
                    ids = ids[p + 2:]
                    continue
                parts.append(embed_table[ids[:p]])                       # :171-179
                parts.append(feat)
                if labels is not None:
                    lparts.append(cur_labels[:p])
                    lparts.append(np.full(feat.shape[0], IGNORE_INDEX, dtype=labels.dtype))
                    cur_labels = cur_labels[p + 1:]
                ids = ids[p + 1:]
            if ids.size > 0:                                             # :182-188
                parts.append(embed_table[ids])
                if labels is not None:
                    lparts.append(cur_labels)
        embeds.append(np.concatenate(parts, axis=0) if parts else np.zeros((0, hdim), embed_table.dtype))
        if labels is not None:
            new_labels.append(np.concatenate(lparts, axis=0))
    lens = [e.shape[0] for e in embeds]
    lmax = max(lens)
    out = np.zeros((B, lmax, hdim), dtype=embed_table.dtype)            # :195-201 right-pad with zeros
    for b, e in enumerate(embeds):
        out[b, :e.shape[0]] = e
    out_labels = None
    if labels is not None:
        out_labels = np.full((B, lmax), IGNORE_INDEX, dtype=labels.dtype)   # :203-209
        for b, l in enumerate(new_labels):
            out_labels[b, :l.shape[0]] = l
    out_mask = attention_mask
    if attention_mask is not None:
        if len(set(lens)) > 1:
            if labels is None:
                raise ValueError("ragged batch without labels: the reference itself fails here (llava_arch.py:211-218 needs labels)")
            out_mask = np.zeros((B, lmax), dtype=attention_mask.dtype)    # :211-219: True x added tokens | old mask | False pad
            for b in range(B):
                added = lens[b] - L
                out_mask[b, :added] = True
                out_mask[b, added:lens[b]] = attention_mask[b]
        else:                                                             # :226-229
            out_mask = np.concatenate([np.ones((B, lmax - L), dtype=attention_mask.dtype), attention_mask], axis=1)
    return out_mask, out, out_labels
