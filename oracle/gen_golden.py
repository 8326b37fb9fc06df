"""Generate tests/golden/*.npz from the REFERENCE ITSELF  --  run in the build container only.

    python oracle/gen_golden.py [/root/reference]

The reference (CircleRadon/TokenPacker) is pure Python and ships no tests or golden vectors, so the known-answer
fixtures are produced by importing its modules by file path and running them on seeded inputs:

  * ``llava/model/multimodal_projector/builder.py``  -> class TokenPacker (imports only torch/numpy)
  * ``llava/patch_divide.py``                        -> class Image_Patch
  * the tiling block has no function boundary (it is pasted inline 9 times); the source range
    ``llava/eval/model_vqa.py:88-123`` is exec'd verbatim in a namespace providing image/image_patch/F/torch
  * the slice assembly ``llava/model/llava_arch.py:141-155`` likewise (namespace provides image_features,
    h_block, w_block, a fake ``self`` whose embed_tokens returns the sep / ret rows)

Nothing here is copied into the repo: the reference source is read and executed from where it lies.
/root/reference does not exist on the GPU box, so tests only ever read the committed fixtures.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import textwrap

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import tokenpacker_oracle as tpo  # noqa: E402
from oracle import hd_oracle as hdo            # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def load_by_path(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def source_range(rel, first, last):
    with open(os.path.join(REF, rel)) as f:
        lines = f.readlines()
    return textwrap.dedent("".join(lines[first - 1:last]))


def ref_projector(builder, params, s, hidden):
    m = builder.TokenPacker(hidden_size=hidden, scale_factor=s)
    sd = {k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in params.items()}
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.eval()


def gen_projector(builder):
    torch.set_num_threads(os.cpu_count())
    # 2 / 3 / 4: the released models; 1 / 6 / 8 / 12 / 24: the other divisors of 24 the constructor accepts (builder.py:51-52)
    for s in (2, 3, 4, 1, 6, 8, 12, 24):
        hidden, n = 128, (2 if s > 1 else 1)
        params = tpo.make_params(hidden, seed=100 + s)
        x0, xm = tpo.make_inputs(n, seed=200 + s)
        with torch.no_grad():
            out = ref_projector(builder, params, s, hidden)((torch.from_numpy(x0), torch.from_numpy(xm))).numpy()
        np.savez(os.path.join(OUT, f"projector_s{s}_h{hidden}.npz"),
                 out=out.astype(np.float32), scale_factor=s, hidden=hidden, n=n,
                 param_seed=100 + s, input_seed=200 + s,
                 x0_probe=x0[0, :3, :5], xm_probe=xm[-1, -3:, -5:],
                 w_probe=params["k_proj_1.0.weight"][:3, :5])
        print("projector", s, out.shape, float(np.abs(out).mean()))
    # bf16-rounded weights and inputs through the fp32 reference: the fixture the GPU parity test uses
    for s in (2, 3, 4):
        hidden, n = 4096, 1
        params = {k: tpo.round_bf16(v) for k, v in tpo.make_params(hidden, seed=0).items()}
        x0, xm = tpo.make_inputs(n, seed=1234 + s)
        x0, xm = tpo.round_bf16(x0), tpo.round_bf16(xm)
        with torch.no_grad():
            out = ref_projector(builder, params, s, hidden)((torch.from_numpy(x0), torch.from_numpy(xm))).numpy()
        np.savez(os.path.join(OUT, f"projector_s{s}_h{hidden}_bf16in.npz"),
                 out_sub=out[0, ::3, ::16].astype(np.float32), row_stride=3, col_stride=16,
                 out_mean=float(out.mean()), out_rms=float(np.sqrt((out.astype(np.float64) ** 2).mean())),
                 scale_factor=s, hidden=hidden, n=n, param_seed=0, input_seed=1234 + s)
        print("projector bf16in", s, out.shape)


def gen_grid(pd):
    rng = np.random.default_rng(7)
    rows = []
    for patch_num in (9, 16, 25):
        ip = pd.Image_Patch(image_size=336, patch_num=patch_num)
        sizes = [(336, 336), (1088, 1088), (224, 1344), (1344, 224), (500, 700), (672, 672), (337, 335), (1, 1),
                 (4000, 300), (300, 4000), (1008, 1008), (1009, 1007), (672, 1008), (50, 5000), (2016, 2016)]
        sizes += [tuple(int(v) for v in rng.integers(32, 2400, size=2)) for _ in range(400)]
        for h, w in sizes:
            hb, wb = ip.calculate(h, w)
            rows.append((h, w, patch_num, hb, wb))
    np.savez(os.path.join(OUT, "hd_grid.npz"), table=np.asarray(rows, dtype=np.int64))
    print("grid rows", len(rows))


def gen_tile(pd):
    src = source_range("llava/eval/model_vqa.py", 88, 123)
    assert src.lstrip().startswith("image = preprocess(image)"), src[:80]
    out = {}
    cases = [(500, 700, 9), (1088, 1088, 9), (336, 336, 9), (224, 1344, 9), (901, 333, 16), (640, 1500, 25)]
    for ci, (h, w, patch_num) in enumerate(cases):
        rng = np.random.default_rng(300 + ci)
        img = rng.standard_normal((3, h, w)).astype(np.float32)
        ns = {"image": torch.from_numpy(img), "preprocess": (lambda t: t),
              "image_patch": pd.Image_Patch(image_size=336, patch_num=patch_num), "F": F, "torch": torch}
        exec(src, ns)
        t = ns["image_tensor"].numpy()
        out[f"case{ci}_meta"] = np.asarray([h, w, patch_num, ns["h_block"], ns["w_block"], 300 + ci], dtype=np.int64)
        out[f"case{ci}_sum"] = t.astype(np.float64).sum(axis=(1, 2, 3))
        out[f"case{ci}_abs"] = np.abs(t.astype(np.float64)).sum(axis=(1, 2, 3))
        out[f"case{ci}_probe"] = t[:, :, ::37, ::41].astype(np.float32)
        print("tile", h, w, patch_num, t.shape)
    out["n_cases"] = np.asarray(len(cases))
    np.savez(os.path.join(OUT, "hd_tile.npz"), **out)


def gen_assemble():
    src = source_range("llava/model/llava_arch.py", 141, 155)
    assert src.lstrip().startswith("image_feature_list = []"), src[:80]
    m, hdim = 3, 4
    grids = [(1, 1), (2, 3), (3, 3), (1, 4), (5, 1)]
    rng = np.random.default_rng(11)
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
    packed, cu = [], [0]
    for b in range(len(grids)):
        ns["batch_idx"] = b
        exec(src, ns)
        seq = ns["cur_image_features"].numpy()
        assert seq.shape[0] == hdo.hd_seq_len(grids[b][0], grids[b][1], m)
        packed.append(seq)
        cu.append(cu[-1] + seq.shape[0])
    np.savez(os.path.join(OUT, "hd_assemble.npz"), feats=feats, grids=np.asarray(grids, dtype=np.int64),
             sep_row=sep_row, ret_row=ret_row, packed=np.concatenate(packed, 0), cu=np.asarray(cu, dtype=np.int64))
    print("assemble", cu)


def gen_splice():
    """Call the reference's own prepare_inputs_labels_for_multimodal (llava_arch.py:100-233) on a stand-in model."""
    import importlib
    import types
    for name, sub in (("llava", "llava"), ("llava.model", "llava/model")):      # bypass the two __init__.py (they import transformers-4.31 symbols)
        mod = types.ModuleType(name)
        mod.__path__ = [os.path.join(REF, sub)]
        sys.modules[name] = mod
    arch = importlib.import_module("llava.model.llava_arch")

    class _Model:
        def __init__(self, table):
            self.table = table

        def embed_tokens(self, ids):
            return self.table[ids]

    class _Tok:
        def convert_tokens_to_ids(self, toks):
            return [{",": 5, "\n": 6}[t] for t in toks]

    class _Fake(arch.LlavaMetaForCausalLM):
        def __init__(self, table, feats):
            self._m, self._f, self.tokenizer = _Model(table), feats, _Tok()
            self.config = types.SimpleNamespace(tune_mm_mlp_adapter=False, mm_use_im_start_end=False)
            self.device = torch.device("cpu")

        def get_model(self):
            return self._m

        def get_vision_tower(self):
            return object()

        def encode_images(self, images):
            return self._f

    rng = np.random.default_rng(77)
    hdim, vocab, m = 16, 64, 4
    table = rng.standard_normal((vocab, hdim)).astype(np.float32)
    out = {"table": table}
    cases = {
        # name: (input_ids, n_images/crop grids, mode, with_labels)
        "equal": ([[1, 2, -200, 3, 4, 7], [9, -200, 8, 10, 11, 12]], None, "pad", True),
        "ragged": ([[1, -200, 3, 4, -200, 7], [9, 13, 8, -200, 11, 12], [20, 21, 22, 23, 24, 25]], None, "pad", True),
        "infer": ([[1, 2, 3, -200, 4]], None, "pad", False),
        "slice": ([[1, -200, 3, 4], [9, 8, -200, 11]], [(2, 2), (1, 1)], "slice", True),
        # tune_mm_mlp_adapter and mm_use_im_start_end (llava_arch.py:162-170): 30 / 31 stand for <im_start> / <im_end>
        "startend": ([[1, 30, -200, 31, 4, 7], [30, -200, 31, 10, 11, 12]], None, "pad", True, True),
        "startend_ragged": ([[1, 30, -200, 31, 30, -200, 31, 7], [9, 13, 8, 30, -200, 31, 11, 12], [20, 21, 22, 23, 24, 25, 26, 27]],
                            None, "pad", True, True),
    }
    for name, case in cases.items():
        ids, grids, mode, with_labels = case[:4]
        start_end = len(case) > 4 and case[4]
        ids_t = torch.tensor(ids)
        n_tok = int((ids_t == -200).sum()) + sum(1 for row in ids if -200 not in row)
        if mode == "slice":
            crops = sum(hdo.n_crops(a, b) for a, b in grids)
            feats = rng.standard_normal((crops, m, hdim)).astype(np.float32)
            hb, wb = [g[0] for g in grids], [g[1] for g in grids]
        else:
            feats = rng.standard_normal((n_tok, m, hdim)).astype(np.float32)
            hb = wb = None
        labels = ids_t.clone() if with_labels else None
        mask = torch.ones_like(ids_t, dtype=torch.bool)
        fake = _Fake(torch.from_numpy(table), torch.from_numpy(feats))
        fake.config = types.SimpleNamespace(tune_mm_mlp_adapter=start_end, mm_use_im_start_end=start_end)
        _, new_mask, _, embeds, new_labels = fake.prepare_inputs_labels_for_multimodal(ids_t, mask, None, labels, object(), mode, hb, wb)
        out[f"{name}_ids"] = np.asarray(ids, dtype=np.int64)
        out[f"{name}_feats"] = feats
        out[f"{name}_embeds"] = embeds.numpy()
        out[f"{name}_mask"] = new_mask.numpy()
        if with_labels:
            out[f"{name}_labels"] = new_labels.numpy()
        if grids is not None:
            out[f"{name}_grids"] = np.asarray(grids, dtype=np.int64)
        print("splice", name, tuple(embeds.shape))
    out["sep_id"], out["ret_id"] = np.asarray(5), np.asarray(6)
    np.savez(os.path.join(OUT, "splice.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    builder = load_by_path("ref_builder", "llava/model/multimodal_projector/builder.py")
    pd = load_by_path("ref_patch_divide", "llava/patch_divide.py")
    gen_projector(builder)
    gen_grid(pd)
    gen_tile(pd)
    gen_assemble()
    gen_splice()
