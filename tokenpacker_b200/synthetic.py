"""Seeded synthetic weights and the algorithmic FLOP / byte model of the path.

numpy-only generators so that the same bits are produced on the build box and the GPU box.  Kept in the product
package (bench.py and examples use it); the oracle carries its own independent copy and a test checks they agree.
"""
from __future__ import annotations

import numpy as np

T, C, CM = 576, 1024, 4096


def param_shapes(hidden: int) -> dict:
    sq, vec = (C, C), (C,)
    d = {"q_proj_1.weight": sq}
    for n in ("k_proj_1", "v_proj_1"):
        d.update({f"{n}.0.weight": (C, CM), f"{n}.0.bias": vec, f"{n}.2.weight": sq, f"{n}.2.bias": vec})
    for n in ("ln_q_1", "ln_k_1", "ln_v_1"):
        d.update({f"{n}.weight": vec, f"{n}.bias": vec})
    d.update({"clip_attn.in_proj_weight": (3 * C, C), "clip_attn.in_proj_bias": (3 * C,),
              "clip_attn.out_proj.weight": sq, "clip_attn.out_proj.bias": vec,
              "mlp.0.weight": (hidden, C), "mlp.0.bias": (hidden,), "mlp.2.weight": (hidden, hidden), "mlp.2.bias": (hidden,)})
    return d


def synthetic_state_dict(hidden: int = 4096, seed: int = 0) -> dict:
    """N(0, 0.02) matrices (the reference's trunc_normal init scale) and N(0, 0.1)-perturbed vectors (LayerNorm weight
    1 + 0.1 z) so that bias and LayerNorm-affine paths carry signal.  Draw order = reference state_dict order."""
    order = ["q_proj_1.weight", "k_proj_1.0.weight", "k_proj_1.0.bias", "k_proj_1.2.weight", "k_proj_1.2.bias",
             "v_proj_1.0.weight", "v_proj_1.0.bias", "v_proj_1.2.weight", "v_proj_1.2.bias",
             "ln_q_1.weight", "ln_q_1.bias", "ln_k_1.weight", "ln_k_1.bias", "ln_v_1.weight", "ln_v_1.bias",
             "clip_attn.in_proj_weight", "clip_attn.in_proj_bias", "clip_attn.out_proj.weight", "clip_attn.out_proj.bias",
             "mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias"]
    shapes = param_shapes(hidden)
    rng = np.random.default_rng(seed)
    out = {}
    for name in order:
        shape = shapes[name]
        if len(shape) == 2:
            out[name] = np.clip(rng.standard_normal(shape).astype(np.float32) * 0.02, -2.0, 2.0)
        else:
            base = 1.0 if (name.startswith("ln_") and name.endswith("weight")) else 0.0
            out[name] = (base + 0.1 * rng.standard_normal(shape).astype(np.float32)).astype(np.float32)
    return out


def flops_per_crop(scale_factor: int, hidden: int = 4096) -> float:
    """2*m*n*k per GEMM in the reference formulation (builder.py:59-83,112-136), MHA in/out projections included."""
    m = (24 // scale_factor) ** 2
    return (2 * (2 * T * CM * C + 2 * T * C * C) + 2 * (2 * T * C * C) + 3 * (2 * m * C * C) + 4 * T * C
            + 2 * m * C * hidden + 2 * m * hidden * hidden)


def bytes_per_crop(scale_factor: int, hidden: int = 4096) -> int:
    """Compulsory HBM bytes per crop, bf16: both feature maps in, the compressed tokens out."""
    m = (24 // scale_factor) ** 2
    return T * (C + CM) * 2 + m * hidden * 2


def weight_bytes(hidden: int = 4096) -> int:
    return 2 * sum(int(np.prod(s)) for s in param_shapes(hidden).values())
