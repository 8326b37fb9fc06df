"""CPU oracle for the TokenPacker projector forward  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithm.  It is the checker
for the CUDA path; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product package (``tokenpacker_b200``) never imports anything under
``oracle/`` and has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors, so
the restatement is pinned against outputs of the reference module itself,
imported by file path in the build container by ``oracle/gen_golden.py``; the
resulting fixtures live in ``tests/golden/`` and ``tests/test_oracle.py``
checks this file against them.

Reference citations are relative to the upstream tree (CircleRadon/TokenPacker):
``llava/model/multimodal_projector/builder.py`` is abbreviated ``builder.py``.
"""
from __future__ import annotations

import math

import numpy as np

try:  # scipy is in the image; keep a slow exact fallback so the oracle never silently degrades
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf, otypes=[np.float64])

RAW_GRID = 24          # builder.py:42   raw_grid=24  (CLIP-ViT-L/14 @336 -> 24x24 patch tokens)
EMBED_DIM = 1024       # builder.py:43
NUM_HEADS = 8          # builder.py:44   1024 // 128
MULTI_DIM = 4096       # builder.py:61,67  hard-coded 4 CLIP layers x 1024
LN_EPS = 1e-6          # builder.py:48


# state_dict keys of the reference module (builder.py:59-83); values are [out, in] like nn.Linear
PARAM_SHAPES = {
    "q_proj_1.weight": (1024, 1024),
    "k_proj_1.0.weight": (1024, 4096), "k_proj_1.0.bias": (1024,),
    "k_proj_1.2.weight": (1024, 1024), "k_proj_1.2.bias": (1024,),
    "v_proj_1.0.weight": (1024, 4096), "v_proj_1.0.bias": (1024,),
    "v_proj_1.2.weight": (1024, 1024), "v_proj_1.2.bias": (1024,),
    "ln_q_1.weight": (1024,), "ln_q_1.bias": (1024,),
    "ln_k_1.weight": (1024,), "ln_k_1.bias": (1024,),
    "ln_v_1.weight": (1024,), "ln_v_1.bias": (1024,),
    "clip_attn.in_proj_weight": (3072, 1024), "clip_attn.in_proj_bias": (3072,),
    "clip_attn.out_proj.weight": (1024, 1024), "clip_attn.out_proj.bias": (1024,),
    # "mlp.0.weight": (H, 1024), "mlp.0.bias": (H,), "mlp.2.weight": (H, H), "mlp.2.bias": (H,)
}


def param_shapes(hidden_size: int) -> dict:
    d = dict(PARAM_SHAPES)
    d["mlp.0.weight"] = (hidden_size, 1024)
    d["mlp.0.bias"] = (hidden_size,)
    d["mlp.2.weight"] = (hidden_size, hidden_size)
    d["mlp.2.bias"] = (hidden_size,)
    return d


def linear(x, w, b=None):
    """nn.Linear: y = x W^T + b."""
    y = x @ w.T
    if b is not None:
        y = y + b
    return y


def gelu_erf(x):
    """nn.GELU() default = exact erf form (builder.py:63,69,81)."""
    return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))


def layer_norm(x, gamma, beta, eps=LN_EPS):
    """nn.LayerNorm(1024, eps=1e-6) over the last dim, biased variance (builder.py:48,73-75)."""
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * gamma + beta


def bilinear_down_weights(n_in: int, n_out: int):
    """1-D weights of F.interpolate(mode='bilinear', align_corners=False) from n_in to n_out.

    ATen's upsample_bilinear2d source index: src = (dst + 0.5) * (n_in / n_out) - 0.5, clamped at 0;
    i0 = floor(src), i1 = min(i0 + 1, n_in - 1), lambda = src - i0.   Used at builder.py:117.
    Returns (i0, i1, w0, w1) as arrays of length n_out.
    """
    scale = n_in / n_out
    dst = np.arange(n_out, dtype=np.float64)
    src = np.maximum((dst + 0.5) * scale - 0.5, 0.0)
    i0 = np.floor(src).astype(np.int64)
    i1 = np.minimum(i0 + 1, n_in - 1)
    w1 = src - i0
    w0 = 1.0 - w1
    return i0, i1, w0, w1


def point_queries(x0, scale_factor: int):
    """builder.py:117-118 — bilinear 24x24 -> gxg point queries (fp32 interpolation in the reference).

    x0: [N, 576, C] -> [N, g*g, C].  Equivalent to a fixed stencil inside each s x s window
    (s=2: mean of the 2x2; s=3: the centre token; s=4: mean of the centre 2x2).
    """
    n, t, c = x0.shape
    g = RAW_GRID // scale_factor
    img = x0.reshape(n, RAW_GRID, RAW_GRID, c)
    i0, i1, w0, w1 = bilinear_down_weights(RAW_GRID, g)
    rows = img[:, i0] * w0[None, :, None, None] + img[:, i1] * w1[None, :, None, None]      # [N,g,24,C]
    out = rows[:, :, i0] * w0[None, None, :, None] + rows[:, :, i1] * w1[None, None, :, None]  # [N,g,g,C]
    return out.reshape(n, g * g, c)


def window_token_index(scale_factor: int):
    """Fine-token indices seen by each query: [M, s*s] into the 576 row-major tokens.

    Restates divide_feature (builder.py:96-105) as used at :122-124: query (hb, wb) attends to fine
    tokens rows hb*s..hb*s+s-1, cols wb*s..wb*s+s-1; key order inside the window is row-major hi*s+wi.
    """
    s = scale_factor
    g = RAW_GRID // s
    hb, wb, hi, wi = np.meshgrid(np.arange(g), np.arange(g), np.arange(s), np.arange(s), indexing="ij")
    idx = (hb * s + hi) * RAW_GRID + (wb * s + wi)
    return idx.reshape(g * g, s * s)


def window_attention(q, k, v, in_w, in_b, out_w, out_b, scale_factor: int, num_heads: int = NUM_HEADS):
    """nn.MultiheadAttention(1024, 8) with L=1 query and S=s*s keys per (query, crop) pair.

    builder.py:77,126-130 -> torch.nn.functional.multi_head_attention_forward slow path:
    packed in-projection, q scaled by 1/sqrt(head_dim), softmax over the s*s keys, out_proj.
    q: [N, M, C]; k, v: [N, 576, C] -> [N, M, C].
    """
    n, m, c = q.shape
    d = c // num_heads
    wq, wk, wv = in_w[:c], in_w[c:2 * c], in_w[2 * c:]
    bq, bk, bv = in_b[:c], in_b[c:2 * c], in_b[2 * c:]
    qp = linear(q, wq, bq) * (1.0 / math.sqrt(d))             # [N,M,C]
    kp = linear(k, wk, bk)                                      # [N,576,C]
    vp = linear(v, wv, bv)
    idx = window_token_index(scale_factor)                      # [M, s*s]
    kw = kp[:, idx].reshape(n, m, -1, num_heads, d)             # [N,M,S,h,d]
    vw = vp[:, idx].reshape(n, m, -1, num_heads, d)
    qh = qp.reshape(n, m, num_heads, d)
    scores = np.einsum("nmhd,nmshd->nmhs", qh, kw)
    scores = scores - scores.max(axis=-1, keepdims=True)
    p = np.exp(scores)
    p = p / p.sum(axis=-1, keepdims=True)
    ctx = np.einsum("nmhs,nmshd->nmhd", p, vw).reshape(n, m, c)
    return linear(ctx, out_w, out_b)


def tokenpacker_forward(params: dict, x0, xm, scale_factor: int, dtype=np.float64):
    """TokenPacker.forward (builder.py:107-137).

    params: reference state_dict as numpy arrays; x0: [N,576,1024]; xm: [N,576,4096].
    Returns [N, (24/s)^2, hidden] in ``dtype`` (float64 default: the checker is more exact than
    either side it referees; pass float32 to mirror the fp32 reference bit-for-bit-ish).
    """
    if RAW_GRID % scale_factor != 0:
        raise ValueError("scale_factor must be divisible by grid size")   # builder.py:51-52 (same message)
    p = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
    x0 = np.asarray(x0, dtype=dtype)
    xm = np.asarray(xm, dtype=dtype)
    # :112-113  key/value feature paths over the multi-level stack
    key = layer_norm(linear(gelu_erf(linear(xm, p["k_proj_1.0.weight"], p["k_proj_1.0.bias"])),
                            p["k_proj_1.2.weight"], p["k_proj_1.2.bias"]),
                     p["ln_k_1.weight"], p["ln_k_1.bias"])
    val = layer_norm(linear(gelu_erf(linear(xm, p["v_proj_1.0.weight"], p["v_proj_1.0.bias"])),
                            p["v_proj_1.2.weight"], p["v_proj_1.2.bias"]),
                     p["ln_v_1.weight"], p["ln_v_1.bias"])
    # :117-120  point queries
    q = point_queries(x0, scale_factor)
    query = layer_norm(linear(q, p["q_proj_1.weight"]), p["ln_q_1.weight"], p["ln_q_1.bias"])
    # :122-134  local-window cross attention
    att = window_attention(query, key, val,
                           p["clip_attn.in_proj_weight"], p["clip_attn.in_proj_bias"],
                           p["clip_attn.out_proj.weight"], p["clip_attn.out_proj.bias"], scale_factor)
    # :136  token refinement MLP
    h = gelu_erf(linear(att, p["mlp.0.weight"], p["mlp.0.bias"]))
    return linear(h, p["mlp.2.weight"], p["mlp.2.bias"])


# ----------------------------------------------------------------------------------------------
# Seeded synthetic weights / inputs shared by tests, smoke() and bench.py.
# numpy-only so that it is identical on the build box and on the GPU box.
# ----------------------------------------------------------------------------------------------

def make_params(hidden_size: int = 4096, seed: int = 0, dtype=np.float32) -> dict:
    """trunc_normal(std=0.02) weights like builder.py:87-94, plus N(0, 0.1) perturbation of every
    1-D parameter so that bias and LayerNorm affine paths are exercised (LN weight = 1 + 0.1 z)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in param_shapes(hidden_size).items():
        if len(shape) == 2:
            w = rng.standard_normal(shape).astype(np.float32) * 0.02
            np.clip(w, -2.0, 2.0, out=w)        # trunc_normal_ default a=-2, b=2 (absolute), a no-op at std .02
            out[name] = w.astype(dtype)
        else:
            z = 0.1 * rng.standard_normal(shape).astype(np.float32)
            base = 1.0 if (name.startswith("ln_") and name.endswith("weight")) else 0.0
            out[name] = (base + z).astype(dtype)
    return out


def make_inputs(n: int, seed: int = 1234, dtype=np.float32):
    rng = np.random.default_rng(seed)
    x0 = rng.standard_normal((n, RAW_GRID * RAW_GRID, EMBED_DIM)).astype(dtype)
    xm = rng.standard_normal((n, RAW_GRID * RAW_GRID, MULTI_DIM)).astype(dtype)
    return x0, xm


def round_bf16(a):
    """Round-to-nearest-even fp32 -> bf16 -> fp32, in numpy (so oracle inputs equal the GPU's bf16 bits)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)


def flops_per_crop(scale_factor: int, hidden: int = 4096) -> float:
    """Algorithmic FLOPs per crop in the reference formulation."""
    t, c, cm = RAW_GRID * RAW_GRID, EMBED_DIM, MULTI_DIM
    m = (RAW_GRID // scale_factor) ** 2
    return (2 * (2 * t * cm * c + 2 * t * c * c) + 2 * (2 * t * c * c) + 3 * (2 * m * c * c)
            + 4 * t * c + 2 * m * c * hidden + 2 * m * hidden * hidden)


def bytes_per_crop(scale_factor: int, hidden: int = 4096) -> int:
    t, c, cm = RAW_GRID * RAW_GRID, EMBED_DIM, MULTI_DIM
    m = (RAW_GRID // scale_factor) ** 2
    return t * (c + cm) * 2 + m * hidden * 2
