"""CPU-only tests of the host-side logic and of the C-ABI boundary (no compute calls)."""
import copy
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import recformer_b200 as rb
from recformer_b200 import _lib
from recformer_b200.engine import FlatParams, _pick_split

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "recformer_b200.h")).read()
    declared = set(re.findall(r"\b(rf_[a-z0-9_]+)\s*\(", header))
    declared -= {"rf_stream_t"}
    assert len(declared) >= 24
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in include/recformer_b200.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    lib = _lib.lib()
    assert lib.rf_version() >= 102
    assert lib.rf_launch_count() == 0


def _header_struct_fields(header, name):
    """Field names of `typedef struct <name> {...} <name>;` in declaration order (comments stripped)."""
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), header, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for piece in decl.split(","):                       # "int M, N, K" -> M, N, K
            fields.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", piece)[-1])
    return fields


@pytest.mark.parametrize("cname,ctype", [("rf_gemm_args", "GemmArgs"), ("rf_attn_args", "AttnArgs"),
                                         ("rf_global_args", "GlobalArgs"), ("rf_embed_args", "EmbedArgs")])
def test_ctypes_structs_mirror_the_header(cname, ctype):
    """The ctypes mirrors in _lib.py must list exactly the header's fields, in order (structs only ever grow at the end;
    a field added on one side only would shift every later argument)."""
    header = open(os.path.join(ROOT, "include", "recformer_b200.h")).read()
    if re.search(r"typedef struct %s \{" % cname, header) is None:
        pytest.skip(f"{cname} is not a struct of this header")
    assert _header_struct_fields(header, cname) == [f[0] for f in getattr(_lib, ctype)._fields_]


def test_argument_validation_errors_without_gpu():
    lib = _lib.lib()
    a = _lib.GemmArgs()
    assert lib.rf_gemm_bf16(ctypes.byref(a), None) == -1
    assert b"empty problem" in lib.rf_last_error()
    assert lib.rf_normalize_rows(None, 0, None, None, 10, 768, None) == -1
    assert lib.rf_cosine_topk(None, None, 1, 1, 768, 0.05, 10, 0, None, None, None, None, None, None) == -1


def test_header_is_plain_c_and_binds_from_a_c_program(tmp_path):
    """The drop-in boundary is a C ABI: include/recformer_b200.h must compile as C (no C++ / torch types) and a C
    program must be able to bind the library with nothing but dlopen — what a non-Python host of the reference would do
    (INTEGRATION.md).  No compute call is made (there is no GPU here): version, argument validation, error string."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    src = tmp_path / "bind.c"
    src.write_text(r"""
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include "recformer_b200.h"
int main(int argc, char** argv) {
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  int (*version)(void) = (int (*)(void))dlsym(h, "rf_version");
  const char* (*last_error)(void) = (const char* (*)(void))dlsym(h, "rf_last_error");
  int (*gemm)(const rf_gemm_args*, rf_stream_t) = (int (*)(const rf_gemm_args*, rf_stream_t))dlsym(h, "rf_gemm_bf16");
  if (!version || !last_error || !gemm) return 3;
  rf_gemm_args a;
  memset(&a, 0, sizeof a);
  int rc = gemm(&a, (rf_stream_t)0);
  printf("%d %d %s\n", version(), rc, last_error());
  return 0;
}
""")
    exe = tmp_path / "bind"
    flags = ["-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", cuda_inc]
    subprocess.run(["gcc", *flags, "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "recformer_b200.h")], check=True)
    subprocess.run(["gcc", *flags, str(src), "-o", str(exe), "-ldl"], check=True)
    out = subprocess.run([str(exe), _lib.LIB_PATH], check=True, capture_output=True, text=True).stdout.split(None, 2)
    assert int(out[0]) >= 102 and int(out[1]) == -1 and "empty problem" in out[2], out


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/librecformer_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_cpu_inputs_are_rejected_not_emulated():
    cfg = rb.RecformerConfig(attention_window=[64], vocab_size=100, num_hidden_layers=1, max_position_embeddings=80)
    m = rb.RecformerModel(cfg)
    ids = torch.zeros(1, 8, dtype=torch.long)
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(input_ids=ids, item_position_ids=ids)


def test_config_defaults_match_reference():
    cfg = rb.RecformerConfig()
    # ref: recformer/models.py:26-38
    assert (cfg.token_type_size, cfg.max_token_num, cfg.max_item_embeddings, cfg.max_attr_num, cfg.max_attr_length,
            cfg.pooler_type, cfg.temp, cfg.mlm_weight, cfg.item_num, cfg.finetune_negative_sample_size) == \
           (4, 2048, 32, 12, 8, "cls", 0.05, 0.1, 0, 0)
    base = rb.RecformerConfig.from_pretrained("allenai/longformer-base-4096")
    assert (base.vocab_size, base.hidden_size, base.num_hidden_layers, base.max_position_embeddings) == (50265, 768, 12, 4098)
    assert base.attention_window == [512] * 12
    with pytest.raises(OSError):
        rb.RecformerConfig.from_pretrained("some/unknown-model")


def test_config_roundtrip(tmp_path):
    cfg = rb.RecformerConfig(attention_window=[64] * 12, max_token_num=1024, max_item_embeddings=51)
    cfg.save_pretrained(str(tmp_path))
    back = rb.RecformerConfig.from_pretrained(str(tmp_path))
    assert back.attention_window == [64] * 12 and back.max_token_num == 1024 and back.max_item_embeddings == 51


def test_constructor_asserts_like_reference():
    with pytest.raises(AssertionError):
        rb.RecformerModel(rb.RecformerConfig(attention_window=[64, 64], num_hidden_layers=3, vocab_size=50,
                                             max_position_embeddings=70))
    with pytest.raises(AssertionError):
        rb.RecformerModel(rb.RecformerConfig(attention_window=63, num_hidden_layers=1, vocab_size=50,
                                             max_position_embeddings=70))
    cfg = rb.RecformerConfig(attention_window=64, num_hidden_layers=2, vocab_size=50, max_position_embeddings=70)
    rb.RecformerModel(cfg)
    assert cfg.attention_window == [64, 64]
    m = rb.RecformerModel(rb.RecformerConfig(attention_window=64, num_hidden_layers=1, vocab_size=50,
                                             max_position_embeddings=70, pooler_type="max"))
    with pytest.raises(NotImplementedError):
        m.pooler(None, torch.zeros(1, 2, 768))


def test_state_dict_keys_match_reference_layout():
    from oracle import recformer_oracle as O
    ocfg = O.OracleConfig(vocab_size=60, num_hidden_layers=2, attention_window=[64, 64], max_position_embeddings=70)
    cfg = rb.RecformerConfig(attention_window=[64, 64], vocab_size=60, num_hidden_layers=2, max_position_embeddings=70,
                             max_item_embeddings=51)
    m = rb.RecformerForSeqRec(cfg)
    want = {k for k, _, _ in O.state_dict_keys(ocfg, "longformer.")} | {"longformer.embeddings.position_ids"}
    assert set(m.state_dict().keys()) == want
    m.init_item_embedding(torch.zeros(5, 768))
    assert "item_embedding.weight" in m.state_dict() and not m.item_embedding.weight.requires_grad
    # padding rows are zero-initialised like HF's _init_weights
    assert m.longformer.embeddings.word_embeddings.weight[1].abs().sum() == 0
    assert m.longformer.get_input_embeddings() is m.longformer.embeddings.word_embeddings


def test_flat_parameter_plan():
    cfg = rb.RecformerConfig(attention_window=[64, 64], vocab_size=60, num_hidden_layers=2, max_position_embeddings=70,
                             max_item_embeddings=51)
    m = rb.RecformerModel(cfg)
    fp: FlatParams = m._engine.params
    assert fp.n_total == sum(p.numel() for p in m.parameters())
    o = fp.offsets
    p = "encoder.layer.1.attention.self."
    assert o[p + "key.weight"] - o[p + "query.weight"] == 768 * 768          # fused QKV weight is contiguous
    assert o[p + "value.weight"] - o[p + "key.weight"] == 768 * 768
    assert o[p + "key.bias"] - o[p + "query.bias"] == 768
    assert fp.n_dense == 2 * (4 * 768 * 768 + 2 * 768 * 3072) and fp.n_dense < fp.n_decay < fp.n_total
    assert all(v % 8 == 0 for v in o.values())


def test_shadow_refresh_skips_the_cast_once_after_the_fused_optimizer(monkeypatch):
    """FusedAdamW writes the bf16 shadow together with the fp32 weights; the next (forced) refresh must not
    cast again, the one after that must, and an in-place update of a dense weight always must."""
    from recformer_b200 import engine as eng
    cfg = rb.RecformerConfig(attention_window=[64], vocab_size=60, num_hidden_layers=1, max_position_embeddings=70,
                             max_item_embeddings=51)
    m = rb.RecformerModel(cfg)
    fp: FlatParams = m._engine.params
    fp.flat = torch.zeros(fp.n_total)                      # host stand-ins: only the bookkeeping is under test
    fp.shadow = torch.zeros(fp.n_dense, dtype=torch.bfloat16)
    casts = []
    monkeypatch.setattr(eng.ops, "cast_bf16", lambda src, dst: casts.append(src.numel()))
    fp.refresh_shadow(force=True)
    assert casts == [fp.n_dense]
    fp.refresh_shadow(force=False)                         # nothing changed: no cast
    assert len(casts) == 1
    fp.mark_shadow_fresh(by_optimizer=True)
    fp.refresh_shadow(force=True)                          # training forward right after optimizer.step(): skipped
    assert len(casts) == 1
    fp.refresh_shadow(force=True)                          # a second forward without a step: forced again
    assert len(casts) == 2
    fp.mark_shadow_fresh(by_optimizer=True)
    with torch.no_grad():
        m.encoder.layer[0].attention.self.query.weight.add_(1.0)     # someone else touched a dense weight
    fp.refresh_shadow(force=False)
    assert len(casts) == 3


def test_adamw_step_scalars_follow_torch_bias_corrections():
    from recformer_b200.optim import FusedAdamW
    cfg = rb.RecformerConfig(attention_window=[64], vocab_size=60, num_hidden_layers=1, max_position_embeddings=70,
                             max_item_embeddings=51)
    opt = FusedAdamW(rb.RecformerModel(cfg), lr=3e-4, betas=(0.9, 0.999))
    opt.step_count = 4                                     # the scalars describe the NEXT step (t = 5)
    lr, bc1, bc2_sqrt, gs = opt.step_scalars(grad_scale=0.125)
    assert lr == 3e-4 and gs == 0.125
    assert abs(bc1 - (1 - 0.9 ** 5)) < 1e-12 and abs(bc2_sqrt - (1 - 0.999 ** 5) ** 0.5) < 1e-12


def test_split_k_heuristic():
    assert _pick_split(2304, 768, 16384) >= 2       # 27 pair-tiles on 74 CTA pairs -> split
    assert _pick_split(768, 768, 128) == 1           # too few K blocks to split
    assert 1 <= _pick_split(3072, 768, 16384) <= 8


def test_tokenizer_layout_matches_reference(goldens):
    g = goldens["tokenizer"]
    cfg = rb.RecformerConfig(max_token_num=1024, max_item_embeddings=51, max_attr_num=3, max_attr_length=32)
    tok = rb.RecformerTokenizer.from_pretrained("allenai/longformer-base-4096", cfg)
    assert tok.batch_encode(copy.deepcopy(g["users"]), encode_item=False, pad_to_max=False) == g["batch"]
    assert tok.batch_encode(copy.deepcopy(g["users"]), encode_item=False, pad_to_max=True) == g["batch_pad_to_max"]
    # SURVEY §8a example
    tok2 = rb.RecformerTokenizer(rb.RecformerConfig(max_token_num=1024, max_item_embeddings=51))
    out = tok2.batch_encode([[([10, 11, 12, 13], [1, 2, 2, 2]), ([20, 21, 22], [1, 2, 2])], [([30, 31], [1, 2])]],
                            encode_item=False)
    assert out["input_ids"] == [[0, 20, 21, 22, 10, 11, 12, 13], [0, 30, 31, 1, 1, 1, 1, 1]]
    assert out["item_position_ids"] == [[0, 1, 1, 1, 2, 2, 2, 2], [0, 1, 1, 50, 50, 50, 50, 50]]
    assert out["global_attention_mask"][0] == [1, 0, 0, 0, 0, 0, 0, 0]
    # attribute text path with a pluggable text tokenizer (ref: tokenization.py:38-61)
    tok3 = rb.RecformerTokenizer(rb.RecformerConfig(max_attr_num=2, max_attr_length=3),
                                 text_tokenizer=lambda s: [ord(c) for c in s])
    ids, tts = tok3.encode_item({"ab": "cde", "f": "g", "dropped": "x"})
    assert ids == [97, 98, 99, 102, 103] and tts == [1, 1, 2, 1, 2]
    t = tok3([[{"a": "b"}], [{"c": "de"}, {"f": "g"}]], return_tensor=True)   # __call__ encodes item dicts
    assert t["input_ids"].dtype == torch.int64 and t["input_ids"].tolist() == [[0, 97, 98, 1, 1, 1], [0, 102, 103, 99, 100, 101]]


def test_ranker_and_topk_ranker_match_reference(goldens):
    for g in goldens["ranker"]:
        s = g["scores"].float()
        got = rb.Ranker([10, 50])(s, g["labels"])
        assert np.allclose(got, g["metrics"], atol=1e-6)
        lab = g["labels"].reshape(-1)
        top = torch.topk(s, 50, dim=-1).values
        tk = rb.TopKRanker([10, 50])(top, s[torch.arange(s.shape[0]), lab])
        assert np.allclose(tk, g["metrics"][:4], atol=1e-6)


def test_ranker_final_batch_of_one_user():
    """A last eval batch of ONE user (len(dataset) % batch_size == 1): the reference squeezes the (1,1) labels to a 0-d
    target, its CE raises and is caught as loss 0.0 while NDCG / Recall / MRR / AUC are still computed (ref:
    utils.py:83-107).  Here the target stays (1,), so the ranking metrics equal the reference's and the loss is the real
    CE instead of the 0.0 placeholder."""
    s = torch.tensor([[0.3, 2.0, -1.0, 0.9, 0.1, 1.5]])
    for lab, rank in ((1, 0), (3, 2), (2, 5)):
        labels = torch.tensor([[lab]])
        got = rb.Ranker([1, 3])(s, labels)
        want = [1 / np.log2(rank + 2) * (rank < 1), float(rank < 1), 1 / np.log2(rank + 2) * (rank < 3), float(rank < 3),
                1 / (rank + 1), 1 - rank / 6]
        assert np.allclose(got[:6], want, atol=1e-6), (lab, got)
        assert np.isclose(got[6], torch.nn.functional.cross_entropy(s, torch.tensor([lab])).item(), atol=1e-6)
        assert rb.Ranker([1, 3])(s, torch.tensor([lab]))[:6] == got[:6]        # (B,) labels are accepted too
        tk = rb.TopKRanker([1, 3])(torch.topk(s, 3).values, s[:, lab])
        assert np.allclose(tk, want[:4], atol=1e-6)


def _imports_of(path):
    """(function name or '<module>', imported module) pairs of one source file, from its AST."""
    import ast
    tree = ast.parse(open(path).read())
    out = []

    def walk(node, scope):
        for child in ast.iter_child_nodes(node):
            s = child.name if isinstance(child, (ast.FunctionDef, ast.AsyncFunctionDef)) else scope
            if isinstance(child, ast.Import):
                out.extend((scope, a.name) for a in child.names)
            elif isinstance(child, ast.ImportFrom):
                out.extend((scope, (child.module or "") + "." + a.name) for a in child.names)
            walk(child, s)

    walk(tree, "<module>")
    return out


def test_only_the_checker_legs_import_the_oracle():
    """The oracle is test infrastructure: the product package and the neutral input generators never import it, and in
    bench.py only the cpu_baseline / --impl reference functions do (the product arm builds its inputs from
    tools/synthetic.py)."""
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    product = glob.glob(os.path.join(root, "recformer_b200", "**", "*.py"), recursive=True) + \
        glob.glob(os.path.join(root, "recformer", "**", "*.py"), recursive=True) + [os.path.join(root, "tools", "synthetic.py")]
    assert len(product) > 10
    for path in product:
        bad = [m for _, m in _imports_of(path) if m.split(".")[0] == "oracle"]
        assert not bad, (path, bad)
    assert not [m for _, m in _imports_of(os.path.join(root, "tools", "synthetic.py")) if m.startswith("recformer")]
    allowed = {"cpu_finetune_step_rate", "cpu_eval_baseline", "cpu_c1_baseline"}
    users = {scope for scope, m in _imports_of(os.path.join(root, "bench.py")) if m.split(".")[0] == "oracle"}
    assert users and users <= allowed, users
    for tool in glob.glob(os.path.join(root, "tools", "*.py")):
        assert not [m for _, m in _imports_of(tool) if m.split(".")[0] == "oracle"], tool
