"""TokenPacker-HD front end: grid selection, crop tiling and slice assembly, host side of the C ABI.

Mirrors the reference seams:
  * ``Image_Patch(image_size=336, patch_num).calculate(h, w)``        llava/patch_divide.py:71-105
  * the inline resize -> pad -> split -> thumbnail block              llava/train/train.py:695-731 (9 pasted copies)
  * the per-image interleaving of crop tokens with ',' / '\\n' rows   llava/model/llava_arch.py:139-155
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import lib, check

BLOCK = 336


def hd_grid(h: int, w: int, patch_num: int = 9, image_size: int = BLOCK):
    hb, wb = C.c_int(0), C.c_int(0)
    check(lib.tp_hd_grid(int(h), int(w), int(patch_num), int(image_size), C.byref(hb), C.byref(wb)), "tp_hd_grid")
    return hb.value, wb.value


class Image_Patch:
    """Same constructor and ``calculate`` contract as patch_divide.py:71-105 (returns the (h_block, w_block) tuple)."""

    def __init__(self, image_size=336, patch_num=9):
        if patch_num not in (9, 16, 25):
            raise NotImplementedError                                    # patch_divide.py:79-80
        if isinstance(image_size, (tuple, list)):
            if image_size[0] != image_size[1]:
                raise NotImplementedError("square crops only (the reference always passes 336)")
            image_size = image_size[0]
        self.image_size = (image_size, image_size)
        self.patch_num = patch_num

    def calculate(self, h, w):
        return hd_grid(h, w, self.patch_num, self.image_size[0])


def hd_fit(h: int, w: int, hb: int, wb: int):
    a, b, c, d = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    check(lib.tp_hd_fit(int(h), int(w), hb, wb, C.byref(a), C.byref(b), C.byref(c), C.byref(d)), "tp_hd_fit")
    return (a.value, b.value), (c.value, d.value)


def n_crops(hb: int, wb: int) -> int:
    return hb * wb + (1 if hb * wb > 1 else 0)


def hd_tile(image: torch.Tensor, patch_num: int = 9):
    """train.py:695-731 on the GPU.  image: float32 CUDA tensor [1,3,h,w] (or [3,h,w]), already normalised.
    Returns (crops [hb*wb(+1), 3, 336, 336] float32, h_block, w_block)."""
    if image.dim() == 4:
        if image.shape[0] != 1:
            raise ValueError("one image per call: [1,3,h,w]")
        image = image[0]
    if image.dim() != 3 or image.shape[0] != 3:
        raise ValueError("image must be [1,3,h,w] or [3,h,w]")
    if not image.is_cuda:
        raise RuntimeError("tokenpacker_b200 has no CPU path: image must be a CUDA tensor")
    image = image.to(torch.float32).contiguous()
    h, w = int(image.shape[1]), int(image.shape[2])
    hb, wb = hd_grid(h, w, patch_num)
    with torch.cuda.device(image.device):
        crops = torch.empty((n_crops(hb, wb), 3, BLOCK, BLOCK), dtype=torch.float32, device=image.device)
        stream = torch.cuda.current_stream(image.device).cuda_stream
        check(lib.tp_hd_tile(image.data_ptr(), h, w, hb, wb, crops.data_ptr(), stream), "tp_hd_tile")
    return crops, hb, wb


_STAGING: dict = {}


def _staging(device, nbytes: int):
    """Pinned host + device byte buffers for the batched tiling plan, one pair per device, grown on demand.  (Calls on one device
    are expected from one stream at a time, like every other entry point of the package.)"""
    key = str(device)
    st = _STAGING.get(key)
    if st is None or st["host"].numel() < nbytes:
        cap = max(1 << 16, 1 << (nbytes - 1).bit_length())
        with torch.cuda.device(device):
            st = {"host": torch.empty(cap, dtype=torch.uint8).pin_memory(), "dev": torch.empty(cap, dtype=torch.uint8, device=device),
                  "event": torch.cuda.Event()}
            st["event"].record(torch.cuda.current_stream(device))
        _STAGING[key] = st
    return st


def hd_tile_batch(images, patch_num: int = 9, _return_launch: bool = False):
    """The tiling block for a whole batch in ONE launch (the collator cats the crops of a batch, train.py:797-800).

    images: sequence of float32 CUDA tensors [3,h,w] or [1,3,h,w] (already normalised), sizes may differ.
    Returns (crops [sum_i n_crops_i, 3, 336, 336] float32 in the reference's order — image by image, grid row-major, thumbnail
    last —, h_block list, w_block list)."""
    imgs = []
    for im in images:
        if im.dim() == 4:
            if im.shape[0] != 1:
                raise ValueError("each image is [3,h,w] or [1,3,h,w]")
            im = im[0]
        if im.dim() != 3 or im.shape[0] != 3:
            raise ValueError("each image is [3,h,w] or [1,3,h,w]")
        if not im.is_cuda:
            raise RuntimeError("tokenpacker_b200 has no CPU path: images must be CUDA tensors")
        imgs.append(im.to(torch.float32).contiguous())
    b = len(imgs)
    if b == 0:
        raise ValueError("empty batch")
    device = imgs[0].device
    hs = (C.c_int64 * b)(*[int(im.shape[1]) for im in imgs])
    ws = (C.c_int64 * b)(*[int(im.shape[2]) for im in imgs])
    ptrs = (C.c_void_p * b)(*[im.data_ptr() for im in imgs])
    hb, wb = (C.c_int * b)(), (C.c_int * b)()
    nc = C.c_int64(0)
    check(lib.tp_hd_tile_batch_plan(hs, ws, ptrs, b, int(patch_num), None, None, hb, wb, C.byref(nc)), "tp_hd_tile_batch_plan")
    # plan tables (image descriptors + crop table) go through ONE pinned staging buffer and ONE asynchronous copy per call
    desc_bytes = C.sizeof(_lib.TpHdImage) * b
    table_off = (desc_bytes + 15) // 16 * 16
    total_bytes = table_off + max(nc.value, 1) * 12
    st = _staging(device, total_bytes)
    st["event"].synchronize()                     # the previous call's copy has left the pinned buffer
    host = st["host"]
    desc = (_lib.TpHdImage * b).from_address(host.data_ptr())
    check(lib.tp_hd_tile_batch_plan(hs, ws, ptrs, b, int(patch_num), desc, C.cast(host.data_ptr() + table_off, C.POINTER(C.c_int32)), hb, wb,
                                    C.byref(nc)), "tp_hd_tile_batch_plan")
    with torch.cuda.device(device):
        dev = st["dev"]
        dev[:total_bytes].copy_(host[:total_bytes], non_blocking=True)
        st["event"].record(torch.cuda.current_stream(device))
        crops = torch.empty((nc.value, 3, BLOCK, BLOCK), dtype=torch.float32, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        check(lib.tp_hd_tile_batch(dev.data_ptr(), dev.data_ptr() + table_off, nc.value, crops.data_ptr(), stream), "tp_hd_tile_batch")
        # the source images must outlive the asynchronous launch: tie them to the stream
        for t in imgs:
            t.record_stream(torch.cuda.current_stream(device))
    if _return_launch:
        # benchmark hook: (device tables, table offset, crop count) so that the kernel can be re-launched and timed on its own
        return crops, list(hb), list(wb), (dev, table_off, nc.value)
    return crops, list(hb), list(wb)


def hd_seq_len(hb: int, wb: int, m: int) -> int:
    return hb * wb * m + hb * (wb - 1) + hb + ((m + 1) if hb * wb > 1 else 0)


@dataclass
class HdPlan:
    n_crops: int
    seg_row_offset: torch.Tensor   # int64 [n_crops]   destination row of each crop's first token
    sep_rows: torch.Tensor         # int64 [n_sep]     rows holding the ',' embedding
    ret_rows: torch.Tensor         # int64 [n_ret]     rows holding the '\n' embedding
    cu_seqlens: torch.Tensor       # int64 [B+1]


def hd_plan(h_block, w_block, tokens_per_crop: int) -> HdPlan:
    hb = [int(v) for v in h_block]
    wb = [int(v) for v in w_block]
    if len(hb) != len(wb):
        raise ValueError("h_block and w_block must have the same length")
    b = len(hb)
    arr = C.c_int * max(b, 1)
    hb_c, wb_c = arr(*hb), arr(*wb)
    nc, ns, nr = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    check(lib.tp_hd_plan(hb_c, wb_c, b, tokens_per_crop, None, None, None, None, C.byref(nc), C.byref(ns), C.byref(nr)), "tp_hd_plan")
    seg = torch.empty(nc.value, dtype=torch.int64)
    sep = torch.empty(ns.value, dtype=torch.int64)
    ret = torch.empty(nr.value, dtype=torch.int64)
    cu = torch.empty(b + 1, dtype=torch.int64)
    p64 = C.POINTER(C.c_int64)
    check(lib.tp_hd_plan(hb_c, wb_c, b, tokens_per_crop, C.cast(seg.data_ptr(), p64), C.cast(sep.data_ptr(), p64),
                         C.cast(ret.data_ptr(), p64), C.cast(cu.data_ptr(), p64), C.byref(nc), C.byref(ns), C.byref(nr)), "tp_hd_plan")
    return HdPlan(nc.value, seg, sep, ret, cu)


_PLAN_CACHE: dict = {}


def hd_plan_device(h_block, w_block, tokens_per_crop: int, device):
    """hd_plan with its index tensors resident on ``device`` (cached per grid signature: serving loops reuse the same few
    grids, and the three small synchronous H2D copies would otherwise sit on the critical path of every call)."""
    key = (tuple(int(v) for v in h_block), tuple(int(v) for v in w_block), int(tokens_per_crop), str(device))
    hit = _PLAN_CACHE.get(key)
    if hit is None:
        plan = hd_plan(h_block, w_block, tokens_per_crop)
        hit = (plan, plan.seg_row_offset.to(device), plan.sep_rows.to(device), plan.ret_rows.to(device))
        if len(_PLAN_CACHE) > 256:
            _PLAN_CACHE.clear()
        _PLAN_CACHE[key] = hit
    return hit


def hd_assemble(feats: torch.Tensor, h_block, w_block, sep_row: torch.Tensor, ret_row: torch.Tensor):
    """llava_arch.py:139-155 for already-projected crop features [sum(crops), M, H] (bf16, CUDA).

    Standalone form of the assembly (used after the multi-GPU all-gather); ``TokenPackerB200.forward_packed`` fuses
    the same scatter into the projector's last GEMM instead.  Returns (packed [sum(L_i), H], cu_seqlens)."""
    if not feats.is_cuda:
        raise RuntimeError("tokenpacker_b200 has no CPU path: feats must be a CUDA tensor")
    m, hdim = int(feats.shape[1]), int(feats.shape[2])
    device = feats.device
    plan, seg, sep_rows, ret_rows = hd_plan_device(h_block, w_block, m, device)
    if plan.n_crops != feats.shape[0]:
        raise ValueError(f"grids describe {plan.n_crops} crops but {feats.shape[0]} were given")
    if torch.is_grad_enabled() and (feats.requires_grad or sep_row.requires_grad or ret_row.requires_grad):
        from .projector import _PackedScatterFunction      # training: differentiable scatter (gradient = one row gather)
        with torch.cuda.device(device):
            out = _PackedScatterFunction.apply(feats.to(torch.bfloat16), sep_row, ret_row, seg, sep_rows, ret_rows, int(plan.cu_seqlens[-1]))
        return (out if feats.dtype == torch.bfloat16 else out.to(feats.dtype)), plan.cu_seqlens
    fb = feats.to(torch.bfloat16).contiguous()
    with torch.cuda.device(device):
        out = torch.empty((int(plan.cu_seqlens[-1]), hdim), dtype=torch.bfloat16, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        check(lib.tp_hd_scatter_crops(fb.data_ptr(), plan.n_crops, m, hdim, seg.data_ptr(), out.data_ptr(), stream), "tp_hd_scatter_crops")
        sep_b = sep_row.to(device=device, dtype=torch.bfloat16).contiguous()
        ret_b = ret_row.to(device=device, dtype=torch.bfloat16).contiguous()
        check(lib.tp_hd_fill_separators(out.data_ptr(), hdim, sep_rows.data_ptr(), sep_rows.numel(), sep_b.data_ptr(),
                                        ret_rows.data_ptr(), ret_rows.numel(), ret_b.data_ptr(), stream), "tp_hd_fill_separators")
    return (out if feats.dtype == torch.bfloat16 else out.to(feats.dtype)), plan.cu_seqlens
