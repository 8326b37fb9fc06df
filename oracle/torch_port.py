"""PyTorch-CPU port of the reference projector forward  --  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference hot path is itself a sequence of PyTorch library calls (builder.py:107-137), and /root/reference does
not exist on the GPU box, so this functional restatement — same ATen ops in the same order, including the five
materialising permute/reshape copies of ``divide_feature`` and ``nn.MultiheadAttention``'s slow path — is what
``bench.py`` times on the host cores as the CPU baseline (``cpu_baseline.kind = "port"``) and as ``--impl reference``.
It is pinned to the same reference-generated fixtures as the numpy oracle (tests/test_oracle.py).
Never imported by the product package.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

RAW_GRID = 24


def _divide(x, kernel_size, token_num, n, c):
    """divide_feature (builder.py:96-105): regroup [T, N, c] into [k*k, (T/k^2)*N, c] windows."""
    h = w = int(token_num ** 0.5)
    k = kernel_size
    x = x.reshape(h // k, k, w, n, c).permute(0, 2, 1, 3, 4)
    x = x.reshape(h // k, w // k, k, k, n, c).permute(0, 1, 3, 2, 4, 5).reshape(h // k, w // k, k * k, n, c)
    return x.permute(2, 0, 1, 3, 4).reshape(k * k, -1, c)


def forward(p: dict, x0: torch.Tensor, xm: torch.Tensor, scale_factor: int, pre_ln=None) -> torch.Tensor:
    """p: reference state_dict (torch tensors, any float dtype); x0 [N,576,1024]; xm [N,576,4096] -> [N,M,H].
    ``pre_ln`` (tests only): applied to the inputs of the three LayerNorms, e.g. a bf16 round trip to model a bf16 module's
    storage of those activations; None = the reference's arithmetic in the tensors' own dtype."""
    if pre_ln is None:
        pre_ln = lambda t: t                                                             # noqa: E731
    if RAW_GRID % scale_factor != 0:
        raise ValueError("scale_factor must be divisible by grid size")
    g = RAW_GRID // scale_factor
    m = g * g

    def two_layer(x, name):
        return F.linear(F.gelu(F.linear(x, p[f"{name}.0.weight"], p[f"{name}.0.bias"])), p[f"{name}.2.weight"], p[f"{name}.2.bias"])

    def ln(x, name):
        return F.layer_norm(x, (1024,), p[f"{name}.weight"], p[f"{name}.bias"], 1e-6)

    key = ln(pre_ln(two_layer(xm, "k_proj_1")), "ln_k_1").permute(1, 0, 2)               # :112
    value = ln(pre_ln(two_layer(xm, "v_proj_1")), "ln_v_1").permute(1, 0, 2)             # :113
    token_num, n, c = key.shape
    q = F.interpolate(x0.reshape(n, RAW_GRID, RAW_GRID, -1).float().permute(0, 3, 1, 2), size=(g, g), mode="bilinear")
    q = q.permute(0, 2, 3, 1).reshape(n, -1, c).to(x0.dtype)                             # :117-118
    query = ln(pre_ln(F.linear(q, p["q_proj_1.weight"])), "ln_q_1").permute(1, 0, 2)     # :120
    rq = _divide(query, 1, m, n, c)                                                      # :122-124
    rk = _divide(key, scale_factor, token_num, n, c)
    rv = _divide(value, scale_factor, token_num, n, c)
    out, _ = F.multi_head_attention_forward(                                             # :126-130 (need_weights default)
        rq, rk, rv, 1024, 8, p["clip_attn.in_proj_weight"], p["clip_attn.in_proj_bias"], None, None, False, 0.0,
        p["clip_attn.out_proj.weight"], p["clip_attn.out_proj.bias"], training=False, need_weights=True)
    x = out.reshape(m, n, -1).permute(1, 0, 2)                                           # :132-134
    return F.linear(F.gelu(F.linear(x, p["mlp.0.weight"], p["mlp.0.bias"])), p["mlp.2.weight"], p["mlp.2.bias"])  # :136
