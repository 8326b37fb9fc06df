"""tokenpacker_b200 — B200-native (sm_100a) implementation of the TokenPacker visual projector hot path.

Drop-in for ``llava/model/multimodal_projector/builder.py`` of CircleRadon/TokenPacker:
``build_vision_projector(config)`` / ``TokenPackerB200.forward((feat, feat_multi))`` keep the reference's
constructor, parameter names and output layout; the arithmetic runs in hand-written tcgen05/TMA CUDA kernels behind
the C ABI declared in ``include/tokenpacker_b200.h``.  There is no CPU fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is not built)
from .projector import TokenPackerB200, TokenPacker, build_vision_projector, IdentityMap
from .hd import Image_Patch, hd_grid, hd_tile, hd_tile_batch, hd_plan, hd_assemble, hd_seq_len
from .splice import splice_multimodal, splice_plan

__all__ = ["TokenPackerB200", "TokenPacker", "build_vision_projector", "IdentityMap", "Image_Patch", "hd_grid", "hd_tile", "hd_tile_batch",
           "hd_plan", "hd_assemble", "hd_seq_len", "splice_multimodal", "splice_plan"]
