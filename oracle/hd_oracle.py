"""CPU oracle for the TokenPacker-HD front end  --  TEST INFRASTRUCTURE ONLY (see tokenpacker_oracle.py).

numpy restatement of
  * the grid selector  ``Image_Patch.calculate``        (llava/patch_divide.py:4-54,57-69,71-105)
  * the tiling block   resize -> pad -> split -> thumb  (llava/train/train.py:695-731; identical copy at
                                                         llava/eval/model_vqa.py:87-123)
  * the slice assembly of per-crop token blocks          (llava/model/llava_arch.py:139-155)

Pinned by ``tests/golden/hd_*.npz`` generated from the reference itself by ``oracle/gen_golden.py``.
"""
from __future__ import annotations

import numpy as np

BLOCK = 336   # train.py:699 block_size

# Candidate (h_block, w_block) tables: data of patch_divide.py:4-54 (order matters: argmax takes the first
# maximum; patches_25 really does list (4,6),(6,4) twice at :52).
_P9 = [(1, 1), (1, 2), (2, 1), (1, 3), (3, 1), (2, 2), (1, 4), (4, 1), (1, 5), (5, 1), (1, 6), (6, 1), (2, 3),
       (3, 2), (1, 7), (7, 1), (4, 2), (2, 4), (1, 8), (8, 1), (3, 3), (1, 9), (9, 1)]
_P16 = _P9 + [(2, 5), (5, 2), (2, 6), (6, 2), (3, 4), (4, 3), (2, 7), (7, 2), (3, 5), (5, 3), (2, 8), (8, 2), (4, 4)]
_P25 = _P16 + [(3, 6), (6, 3), (2, 9), (9, 2), (4, 5), (5, 4), (2, 10), (10, 2), (3, 7), (7, 3), (11, 2), (2, 11),
               (4, 6), (6, 4), (12, 2), (2, 12), (3, 8), (8, 3), (4, 6), (6, 4), (5, 5)]
GRID_TABLES = {9: _P9, 16: _P16, 25: _P25}

f32 = np.float32


def hd_grid(h: int, w: int, patch_num: int = 9, image_size: int = BLOCK):
    """Image_Patch(image_size, patch_num).calculate(h, w)  (patch_divide.py:96-105) in float32 arithmetic,
    op for op: int64 operands are converted to float32 at each mixed op exactly where torch promotes."""
    if patch_num not in GRID_TABLES:
        raise NotImplementedError                                    # patch_divide.py:79-80
    table = GRID_TABLES[patch_num]
    ph = np.array([p[0] * image_size for p in table], dtype=np.int64)   # patches[:, 2]
    pw = np.array([p[1] * image_size for p in table], dtype=np.int64)   # patches[:, 3]
    area1 = (ph * pw)                                                    # box_area, int64          (:94)
    # :98-99  ratio = patches[:, 2:] / input_box[:, 2:]; min over the two
    ratio = np.minimum(ph.astype(f32) / f32(h), pw.astype(f32) / f32(w))
    # :100    torch.round is round-half-to-even, as is np.round
    score = np.round(f32(h) * ratio) * np.round(f32(w) * ratio) / area1.astype(f32)
    # :101    box_iou(patches, areas, input_box * 1.4)   (:57-69)
    bh = f32(h) * f32(1.4)
    bw = f32(w) * f32(1.4)
    area2 = (bh - f32(0)) * (bw - f32(0))
    wh0 = np.maximum(np.minimum(ph.astype(f32), bh) - f32(0), f32(0))
    wh1 = np.maximum(np.minimum(pw.astype(f32), bw) - f32(0), f32(0))
    inter = wh0 * wh1
    union = area1.astype(f32) + area2 - inter
    iou = inter / (union + f32(1e-5))
    score = score + iou * f32(0.1)                                       # :103
    return table[int(np.argmax(score))]                                  # :104-105


def _linear_taps(n_in: int, n_out: int):
    """ATen upsample_bilinear2d (align_corners=False, no antialias) taps in float32.

    scale = n_in / n_out (float32); src = fma(scale, dst + 0.5, -0.5), clamped below at 0;
    i0 = int(src); i1 = i0 + (i0 < n_in - 1); w1 = src - i0; w0 = 1 - w1.

    The source coordinate is ONE fused multiply-add in ATen's CPU kernel as built (x86 vector code), not a rounded product
    followed by a rounded subtraction: at ~1000-pixel extents one float32 ulp of the coordinate is 6e-5 of a pixel, so the two
    differ by up to ~4e-5 in the output for O(1) pixel values.  Found by tests/test_reference_live.py on random sizes (the six
    fixture cases happen to agree either way to 2e-6); emulated here exactly via float64 (24x24-bit product is exact).
    """
    scale = f32(n_in) / f32(n_out)
    dst = np.arange(n_out, dtype=f32)
    src = (np.float64(scale) * (dst + f32(0.5)).astype(np.float64) - 0.5).astype(f32)
    src = np.maximum(src, f32(0)).astype(f32)
    i0 = src.astype(np.int64)
    i1 = np.minimum(i0 + 1, n_in - 1)
    w1 = (src - i0.astype(f32)).astype(f32)
    w0 = (f32(1) - w1).astype(f32)
    return i0, i1, w0, w1


def resize_bilinear(img, h_out: int, w_out: int):
    """F.interpolate(img[1,3,h,w], size=(h_out,w_out), mode='bilinear') in float32 (train.py:709,727)."""
    img = np.asarray(img, dtype=f32)
    _, _, h, w = img.shape
    y0, y1, wy0, wy1 = _linear_taps(h, h_out)
    x0, x1, wx0, wx1 = _linear_taps(w, w_out)
    top = img[:, :, y0][:, :, :, x0] * wx0 + img[:, :, y0][:, :, :, x1] * wx1
    bot = img[:, :, y1][:, :, :, x0] * wx0 + img[:, :, y1][:, :, :, x1] * wx1
    return (top * wy0[:, None] + bot * wy1[:, None]).astype(f32)


def _fit(h, w, hb, wb):
    """train.py:701-708 / :719-726 — target size with Python round() (banker's) and the min clamp."""
    h_ratio = BLOCK * hb / h
    w_ratio = BLOCK * wb / w
    if h_ratio <= w_ratio:
        return BLOCK * hb, min(BLOCK * wb, round(w * h_ratio))
    return min(BLOCK * hb, round(h * w_ratio)), BLOCK * wb


def hd_tile(image, patch_num: int = 9):
    """train.py:695-731.  image: float32 [1,3,h,w] (already ToTensor+Normalize'd).
    Returns (crops [hb*wb(+1), 3, 336, 336] float32, hb, wb).  Crop order: row-major grid, then thumbnail.
    NB the thumbnail is a resize of the zero-PADDED canvas (``image`` was rebound at :710), fitted to the
    ORIGINAL aspect ratio — restated faithfully."""
    image = np.asarray(image, dtype=f32)
    h, w = image.shape[-2:]
    hb, wb = hd_grid(h, w, patch_num)
    h_, w_ = _fit(h, w, hb, wb)
    canvas = np.zeros((1, 3, BLOCK * hb, BLOCK * wb), dtype=f32)
    canvas[:, :, :h_, :w_] = resize_bilinear(image, h_, w_)
    crops = [canvas[:, :, BLOCK * i:BLOCK * (i + 1), BLOCK * j:BLOCK * (j + 1)]
             for i in range(hb) for j in range(wb)]
    if len(crops) > 1:
        th, tw = _fit(h, w, 1, 1)
        thumb = np.zeros((1, 3, BLOCK, BLOCK), dtype=f32)
        thumb[:, :, :th, :tw] = resize_bilinear(canvas, th, tw)
        crops.append(thumb)
    return np.concatenate(crops, axis=0), hb, wb


def n_crops(hb: int, wb: int) -> int:
    return hb * wb + (1 if hb * wb > 1 else 0)


def hd_seq_len(hb: int, wb: int, m: int) -> int:
    """Length of one image's assembled sequence (llava_arch.py:141-152)."""
    return hb * wb * m + hb * (wb - 1) + hb + ((m + 1) if hb * wb > 1 else 0)


def hd_assemble(feats, h_block, w_block, sep_row, ret_row):
    """llava_arch.py:139-155 for a batch of images.

    feats: [sum(crops), M, H]; per image: for each grid row: crop tokens, a ',' embedding row between columns,
    a '\\n' row at the end of the grid row; if more than one crop, thumbnail tokens + '\\n'.
    Returns (packed [sum(L_i), H], cu_seqlens [B+1])."""
    feats = np.asarray(feats)
    out, cu, idx = [], [0], 0
    for hb, wb in zip(h_block, w_block):
        parts = []
        for _h in range(hb):
            for _w in range(wb):
                parts.append(feats[idx]); idx += 1
                if _w < wb - 1:
                    parts.append(sep_row[None])
            parts.append(ret_row[None])
        if hb * wb > 1:
            parts.append(feats[idx]); idx += 1
            parts.append(ret_row[None])
        seq = np.concatenate(parts, axis=0)
        out.append(seq)
        cu.append(cu[-1] + seq.shape[0])
    return np.concatenate(out, axis=0), np.asarray(cu, dtype=np.int64)
