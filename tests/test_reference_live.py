"""Randomised comparison against the REFERENCE ITSELF, imported from /root/reference when it is mounted (the build container);
skipped elsewhere (the GPU box has no reference: GPU tests only ever use the committed fixtures).  Widens the pin of the oracle
and of the product's host logic beyond the fixed fixture cases: fresh seeds every run of this file would defeat reproducibility,
so the seeds are fixed but different from the fixture seeds.  Nothing is copied: modules are loaded from where they lie."""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

REF = os.environ.get("TOKENPACKER_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "llava")), reason="reference tree not mounted")


def _by_path(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def ref_builder():
    return _by_path("ref_builder_live", "llava/model/multimodal_projector/builder.py")


@pytest.fixture(scope="module")
def ref_patch_divide():
    return _by_path("ref_patch_divide_live", "llava/patch_divide.py")


@pytest.fixture(scope="module")
def ref_arch():
    for name, sub in (("llava", "llava"), ("llava.model", "llava/model")):      # bypass the two __init__.py (transformers-4.31 imports)
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = [os.path.join(REF, sub)]
            sys.modules[name] = mod
    return importlib.import_module("llava.model.llava_arch")


@pytest.mark.parametrize("s,hidden,seed", [(2, 64, 901), (3, 96, 902), (4, 160, 903), (6, 32, 904), (12, 64, 905)])
def test_oracle_vs_reference_module(ref_builder, s, hidden, seed):
    """Fresh weights (every 1-D parameter perturbed so LayerNorm / bias paths matter), odd hidden sizes, N=2."""
    import torch
    from oracle import tokenpacker_oracle as tpo
    from oracle import torch_port
    params = tpo.make_params(hidden, seed=seed)
    x0, xm = tpo.make_inputs(2, seed=seed + 1000)
    m = ref_builder.TokenPacker(hidden_size=hidden, scale_factor=s)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    with torch.no_grad():
        ref = m.eval()((torch.from_numpy(x0), torch.from_numpy(xm))).numpy()
        port = torch_port.forward({k: torch.from_numpy(v) for k, v in params.items()}, torch.from_numpy(x0), torch.from_numpy(xm), s).numpy()
    out = tpo.tokenpacker_forward(params, x0, xm, s)
    assert np.abs(out - ref).max() < 5e-6
    assert np.abs(port - ref).max() < 1e-6


def test_grid_selector_vs_reference_random(ref_patch_divide):
    """Product (C ABI, host arithmetic) and oracle vs Image_Patch.calculate on 600 fresh sizes incl. extreme aspect ratios."""
    from oracle import hd_oracle as hdo
    from tokenpacker_b200 import hd_grid
    rng = np.random.default_rng(4242)
    for patch_num in (9, 16, 25):
        ip = ref_patch_divide.Image_Patch(image_size=336, patch_num=patch_num)
        sizes = [tuple(int(v) for v in rng.integers(16, 3200, size=2)) for _ in range(170)]
        sizes += [(int(rng.integers(16, 200)), int(rng.integers(2000, 6000))) for _ in range(15)]
        sizes += [(int(rng.integers(2000, 6000)), int(rng.integers(16, 200))) for _ in range(15)]
        for h, w in sizes:
            want = tuple(int(v) for v in ip.calculate(h, w))
            assert hd_grid(h, w, patch_num) == want, (h, w, patch_num)
            assert tuple(hdo.hd_grid(h, w, patch_num)) == want, (h, w, patch_num)


def _fake_model(arch, table, feats, start_end):
    import torch

    class _Model:
        def embed_tokens(self, ids):
            return table[ids]

    class _Tok:
        def convert_tokens_to_ids(self, toks):
            return [{",": 5, "\n": 6}[t] for t in toks]

    class _Fake(arch.LlavaMetaForCausalLM):
        def __init__(self):
            self._m, self.tokenizer = _Model(), _Tok()
            self.config = types.SimpleNamespace(tune_mm_mlp_adapter=start_end, mm_use_im_start_end=start_end)
            self.device = torch.device("cpu")

        def get_model(self):
            return self._m

        def get_vision_tower(self):
            return object()

        def encode_images(self, images):
            return feats

    return _Fake()


@pytest.mark.parametrize("start_end", [False, True])
def test_splice_vs_reference_method_random(ref_arch, start_end):
    """Random batches through the reference's prepare_inputs_labels_for_multimodal vs the oracle AND the product's host planner
    (llava_arch.py:100-233, both mm_use_im_start_end branches, 'pad' and 'slice' modes, ragged and image-free samples)."""
    import torch
    from oracle import hd_oracle as hdo
    from oracle import splice_oracle as spo
    from tokenpacker_b200 import splice_plan
    rng = np.random.default_rng(77 if start_end else 78)
    hdim, vocab, m = 8, 40, 3
    table = rng.standard_normal((vocab, hdim)).astype(np.float32)
    for trial in range(40):
        B, L = int(rng.integers(1, 4)), int(rng.integers(6, 12))
        slice_mode = (not start_end) and trial % 2 == 1
        ids = rng.integers(7, vocab, size=(B, L))
        n_img = []
        for b in range(B):
            k = 1 if slice_mode else int(rng.integers(0, 3))
            if start_end:
                # <im_start> IMAGE <im_end> triples (30 / 31 stand-ins), never at position 0 (upstream always has a BOS first)
                pos = sorted(rng.choice(np.arange(2, L - 1, 3), size=min(k, (L - 3) // 3), replace=False).tolist())
                for p in pos:
                    ids[b, p - 1], ids[b, p], ids[b, p + 1] = 30, -200, 31
                n_img.append(len(pos))
            else:
                pos = sorted(rng.choice(L, size=k, replace=False).tolist())
                ids[b, pos] = -200
                n_img.append(k)
        labels = ids.copy()
        mask = np.ones_like(ids, dtype=bool)
        if slice_mode:
            grids = [(int(rng.integers(1, 4)), int(rng.integers(1, 4))) for _ in range(B)]
            crops = sum(hdo.n_crops(a, b) for a, b in grids)
            feats = rng.standard_normal((crops, m, hdim)).astype(np.float32)
            hb, wb = [g[0] for g in grids], [g[1] for g in grids]
            packed, cu = hdo.hd_assemble(feats, hb, wb, table[5], table[6])
            seqs = [packed[cu[i]:cu[i + 1]] for i in range(B)]
            mode = "slice"
        else:
            n_seq = sum(max(k, 1) for k in n_img)          # an image-free sample still consumes one (llava_arch.py:121-134)
            feats = rng.standard_normal((n_seq, m, hdim)).astype(np.float32)
            seqs = [feats[i] for i in range(n_seq)]
            hb = wb = None
            mode = "pad"
        fake = _fake_model(ref_arch, torch.from_numpy(table), torch.from_numpy(feats), start_end)
        _, ref_mask, _, ref_embeds, ref_labels = fake.prepare_inputs_labels_for_multimodal(
            torch.from_numpy(ids), torch.from_numpy(mask), None, torch.from_numpy(labels), object(), mode, hb, wb)
        o_mask, o_embeds, o_labels = spo.splice(ids, mask, labels, seqs, table, im_start_end=start_end)
        np.testing.assert_array_equal(o_embeds, ref_embeds.numpy())
        np.testing.assert_array_equal(o_labels, ref_labels.numpy())
        np.testing.assert_array_equal(o_mask, ref_mask.numpy())
        visual = np.concatenate(seqs, axis=0)
        cu_seq = np.concatenate([[0], np.cumsum([q.shape[0] for q in seqs])])
        plan = splice_plan(ids, cu_seq, labels, mask, im_start_end=start_end)
        rows = np.zeros((plan.src_index.shape[0], hdim), dtype=np.float32)
        src = plan.src_index
        rows[src >= 0] = table[src[src >= 0]]
        rows[src <= -2] = visual[-src[src <= -2] - 2]
        np.testing.assert_array_equal(rows.reshape(B, plan.lmax, hdim), ref_embeds.numpy())
        np.testing.assert_array_equal(plan.labels, ref_labels.numpy())
        np.testing.assert_array_equal(plan.attention_mask, ref_mask.numpy())


def test_tiling_block_vs_reference_source_random(ref_patch_divide):
    """The resize -> pad -> split -> thumbnail block has no function boundary upstream (pasted inline 9 times); the source range
    eval/model_vqa.py:88-123 is exec'd where it lies and compared with the oracle restatement on fresh image sizes."""
    import textwrap
    import torch
    import torch.nn.functional as F
    from oracle import hd_oracle as hdo
    with open(os.path.join(REF, "llava/eval/model_vqa.py")) as f:
        src = textwrap.dedent("".join(f.readlines()[87:123]))
    assert src.lstrip().startswith("image = preprocess(image)")
    rng = np.random.default_rng(515)
    for trial in range(18):
        patch_num = (9, 16, 25)[trial % 3]
        h, w = (int(v) for v in rng.integers(40, 1500, size=2))
        img = rng.standard_normal((3, h, w)).astype(np.float32)
        ns = {"image": torch.from_numpy(img), "preprocess": (lambda t: t),
              "image_patch": ref_patch_divide.Image_Patch(image_size=336, patch_num=patch_num), "F": F, "torch": torch}
        exec(src, ns)
        want = ns["image_tensor"].numpy()
        crops, hb, wb = hdo.hd_tile(img[None], patch_num)
        assert (hb, wb) == (int(ns["h_block"]), int(ns["w_block"]))
        assert crops.shape == want.shape, (crops.shape, want.shape)
        assert np.abs(crops - want).max() <= 2e-6, (h, w, patch_num, float(np.abs(crops - want).max()))


@pytest.mark.parametrize("s,hidden,seed", [(2, 64, 911), (3, 32, 912), (4, 96, 913), (8, 32, 914)])
def test_gradient_oracle_vs_reference_autograd(ref_builder, s, hidden, seed):
    """tests/test_backward_gpu.py uses autograd over oracle/torch_port.py as the gradient oracle: pin THAT to autograd through the
    reference module itself (fp32, CPU), every parameter."""
    import torch
    from oracle import tokenpacker_oracle as tpo
    from oracle import torch_port
    params = tpo.make_params(hidden, seed=seed)
    x0, xm = tpo.make_inputs(2, seed=seed + 1000)
    x0, xm = torch.from_numpy(x0), torch.from_numpy(xm)
    gw = torch.from_numpy(np.random.default_rng(seed).standard_normal((2, (24 // s) ** 2, hidden)).astype(np.float32))
    m = ref_builder.TokenPacker(hidden_size=hidden, scale_factor=s)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    (m((x0, xm)) * gw).sum().backward()
    p = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in params.items()}
    (torch_port.forward(p, x0, xm, s) * gw).sum().backward()
    for name, ref_param in m.named_parameters():
        g_ref, g = ref_param.grad, p[name].grad
        scale = float(g_ref.abs().max()) + 1e-12
        assert float((g - g_ref).abs().max()) <= 2e-5 * scale + 1e-7, (name, float((g - g_ref).abs().max()), scale)


def test_slice_assembly_vs_reference_source_random():
    """llava_arch.py:141-155 exec'd where it lies on random grids vs the oracle and the product's host plan (tp_hd_plan)."""
    import textwrap
    import torch
    from oracle import hd_oracle as hdo
    from tokenpacker_b200 import hd_plan
    with open(os.path.join(REF, "llava/model/llava_arch.py")) as f:
        src = textwrap.dedent("".join(f.readlines()[140:155]))
    assert src.lstrip().startswith("image_feature_list = []")
    rng = np.random.default_rng(606)
    for trial in range(20):
        m, hdim = int(rng.integers(1, 6)), 4
        grids = [(int(rng.integers(1, 6)), int(rng.integers(1, 6))) for _ in range(int(rng.integers(1, 6)))]
        sep_row = rng.standard_normal(hdim).astype(np.float32)
        ret_row = rng.standard_normal(hdim).astype(np.float32)
        total = sum(hdo.n_crops(a, b) for a, b in grids)
        feats = rng.standard_normal((total, m, hdim)).astype(np.float32)

        class _Model:
            def embed_tokens(self, tok):
                return torch.from_numpy(sep_row if int(tok[0]) == 0 else ret_row)[None]

        class _Self:
            def get_model(self):
                return _Model()

        ns = {"image_features": torch.from_numpy(feats), "h_block": [g[0] for g in grids], "w_block": [g[1] for g in grids],
              "self": _Self(), "sep": torch.tensor([0]), "ret": torch.tensor([1]), "torch": torch, "cur_image_idx": 0}
        want = []
        for b in range(len(grids)):
            ns["batch_idx"] = b
            exec(src, ns)
            want.append(ns["cur_image_features"].numpy())
        want_cu = np.concatenate([[0], np.cumsum([q.shape[0] for q in want])])
        want = np.concatenate(want, axis=0)
        hb, wb = [g[0] for g in grids], [g[1] for g in grids]
        packed, cu = hdo.hd_assemble(feats, hb, wb, sep_row, ret_row)
        np.testing.assert_array_equal(packed, want)
        np.testing.assert_array_equal(cu, want_cu)
        plan = hd_plan(hb, wb, m)
        np.testing.assert_array_equal(plan.cu_seqlens.numpy(), want_cu)
        rebuilt = np.full_like(want, np.nan)
        for c, r0 in enumerate(plan.seg_row_offset.tolist()):
            rebuilt[r0:r0 + m] = feats[c]
        rebuilt[plan.sep_rows.numpy()] = sep_row
        rebuilt[plan.ret_rows.numpy()] = ret_row
        np.testing.assert_array_equal(rebuilt, want)
