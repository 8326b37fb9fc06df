"""Backward of the projector (tp_forward_train / tp_backward through autograd) against PyTorch autograd over the oracle's
torch port (the reference's op sequence, fp32) on identical bf16-rounded weights and inputs.
Tolerance: every parameter gradient within 3e-2 relative RMS error (bf16 activations and bf16 gradient storage through a
chain of ~10 GEMMs; the forward's own gate is 3e-3)."""
import numpy as np
import pytest
import torch

from oracle import tokenpacker_oracle as tpo
from oracle import torch_port

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("s,hidden,n", [(2, 256, 2), (3, 128, 3), (4, 256, 4), (6, 256, 2), (24, 256, 3)])
def test_parameter_gradients_match_autograd_of_oracle(s, hidden, n):
    from tokenpacker_b200 import TokenPackerB200
    params = {k: tpo.round_bf16(v) for k, v in tpo.make_params(hidden, seed=21 + s).items()}
    m = TokenPackerB200(hidden_size=hidden, scale_factor=s)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    m = m.to("cuda", torch.bfloat16).train()
    x0, xm = tpo.make_inputs(n, seed=31 + s)
    x0 = torch.from_numpy(tpo.round_bf16(x0)).cuda()
    xm = torch.from_numpy(tpo.round_bf16(xm)).cuda()
    gen = torch.Generator(device="cuda").manual_seed(5)
    gw = torch.randn(n, (24 // s) ** 2, hidden, device="cuda", generator=gen)

    out = m((x0.bfloat16(), xm.bfloat16()))
    (out.float() * gw.bfloat16().float()).sum().backward()

    ref_p = {k: torch.from_numpy(v).cuda().requires_grad_(True) for k, v in params.items()}
    ref_out = torch_port.forward(ref_p, x0, xm, s)
    (ref_out * gw.bfloat16().float()).sum().backward()

    assert (out.float() - ref_out).pow(2).mean().sqrt() / ref_out.pow(2).mean().sqrt() < 3e-3
    worst = {}
    for name, p in m.named_parameters():
        g, r = p.grad.float(), ref_p[name].grad
        assert g.shape == r.shape and torch.isfinite(g).all(), name
        worst[name] = (float((g - r).pow(2).mean().sqrt()), float(r.pow(2).mean().sqrt()))
    # absolute floor: the softmax is invariant to a constant shift of all keys of a window, so the gradients of ln_k_1.bias and
    # of the k slice of in_proj_bias are analytically ZERO (reference: ~1e-8) and k_proj_1.2.bias is nearly so; there the
    # comparison is against rounding noise, bounded at 1 % of the k branch's own first-layer bias gradient
    floor = 1e-2 * worst["k_proj_1.0.bias"][1]
    bad = {k: (e, r) for k, (e, r) in worst.items() if e > 3e-2 * r + floor}
    assert not bad, (bad, worst)
    assert max(e / r for k, (e, r) in worst.items() if r > 30 * floor) < 1.5e-2      # every non-degenerate gradient: < 1.5 % rel-RMS


def test_training_forward_equals_inference_forward_closely_and_step_changes_output():
    from tokenpacker_b200 import TokenPackerB200
    s, hidden, n = 2, 256, 2
    params = {k: tpo.round_bf16(v) for k, v in tpo.make_params(hidden, seed=2).items()}
    m = TokenPackerB200(hidden_size=hidden, scale_factor=s)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    m = m.to("cuda", torch.bfloat16)
    g = torch.Generator(device="cuda").manual_seed(1)
    x0 = torch.randn(n, 576, 1024, device="cuda", generator=g).bfloat16()
    xm = torch.randn(n, 576, 4096, device="cuda", generator=g).bfloat16()
    with torch.no_grad():
        inf = m((x0, xm))
    tr = m((x0, xm))
    assert tr.requires_grad
    # training forward does not fold out_proj into mlp.0 and applies GELU as its own pass: same math, different roundings
    assert (tr.float() - inf.float()).abs().max().item() < 1e-2
    opt = torch.optim.SGD(m.parameters(), lr=10.0)        # large enough that the update survives the bf16 rounding of the parameters
    tr.float().pow(2).mean().backward()
    opt.step()
    with torch.no_grad():
        after = m((x0, xm))                      # weight cache must notice the in-place update
    assert after.float().pow(2).mean() < inf.float().pow(2).mean()


@pytest.mark.parametrize("hidden", [256, 4096])
def test_train_forward_in_place_weights_equal_packed_copies(hidden):
    """tp_forward_train reads the untransformed weight matrices either from the live parameters (w given + tp_pack_weights_train)
    or from their copies in a full tp_pack_weights buffer (w NULL): same kernels on the same bits, so the outputs are identical."""
    import ctypes as C
    from tokenpacker_b200 import TokenPackerB200, _lib
    lib = _lib.lib
    s, n = 2, 3
    params = {k: tpo.round_bf16(v) for k, v in tpo.make_params(hidden, seed=77).items()}
    m = TokenPackerB200(hidden_size=hidden, scale_factor=s)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    m = m.to("cuda", torch.bfloat16)
    x0, xm = tpo.make_inputs(n, seed=78)
    x0 = torch.from_numpy(tpo.round_bf16(x0)).cuda().bfloat16().contiguous()
    xm = torch.from_numpy(tpo.round_bf16(xm)).cuda().bfloat16().contiguous()
    bf = [p.detach().contiguous() for p in m._raw_params()]
    w = _lib.TpWeights(*[t.data_ptr() for t in bf])
    stream = torch.cuda.current_stream().cuda_stream
    pbytes = lib.tp_packed_bytes(hidden)
    sbytes = lib.tp_train_saved_bytes(n, s, hidden)
    outs = []
    for in_place in (False, True):
        packed = torch.zeros(pbytes, dtype=torch.uint8, device="cuda")
        saved = torch.empty(sbytes, dtype=torch.uint8, device="cuda")
        out = torch.empty(n, m.num_queries, hidden, dtype=torch.bfloat16, device="cuda")
        pack = lib.tp_pack_weights_train if in_place else lib.tp_pack_weights
        _lib.check(pack(C.byref(w), hidden, packed.data_ptr(), pbytes, stream), "pack")
        _lib.check(lib.tp_forward_train(C.byref(w) if in_place else None, packed.data_ptr(), x0.data_ptr(), xm.data_ptr(), n, 576 * 1024, 576 * 4096,
                                        s, hidden, out.data_ptr(), saved.data_ptr(), sbytes, stream), "tp_forward_train")
        outs.append(out)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    with torch.no_grad():
        inf = m.eval()((x0, xm))
    assert (outs[1].float() - inf.float()).abs().max().item() < 1e-2
