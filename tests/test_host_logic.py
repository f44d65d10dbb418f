"""CPU tests of the host side: string packing, slot bounds, frame assembly, the C-ABI library
(loads, exports every symbol include/pfz.h declares), plugin contract, loud failure without a GPU."""
import os
import re

import numpy as np
import pytest
import torch

import polyfuzz_b200
from polyfuzz_b200 import _lib
from polyfuzz_b200.strings import pack_utf32, ngram_slot_bounds
from polyfuzz_b200.matchers._utils import assemble_matches, clip_top_n
from oracle import tfidf as otfidf
from oracle.assemble import assemble as oracle_assemble

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_pack_utf32_roundtrip():
    s = ["apple", "", "İstanbul", "a\U0001F600b", "中文 K"]
    blob, offs = pack_utf32(s)
    assert offs.tolist() == [0, 5, 5, 13, 16, 20]
    for i, x in enumerate(s):
        assert "".join(chr(c) for c in blob[offs[i]:offs[i + 1]]) == x
    blob, offs = pack_utf32([])
    assert blob.size == 0 and offs.tolist() == [0]
    with pytest.raises(TypeError):
        pack_utf32(["a", 3])


def test_pack_strings_c_extension_equals_python_path():
    from polyfuzz_b200 import strings
    from polyfuzz_b200 import synth
    assert strings._hostpack() is not None, "polyfuzz_b200/_pfz_hostpack*.so missing: run python -m polyfuzz_b200.build"
    for lst in (synth.company_names(500, seed=1), synth.titles(500, seed=2) + ["", "İK", "a\U0001F600b"], ("tuple", "too"), []):
        b, o = pack_utf32(list(lst))
        b2, o2, _ = strings.pack_strings(lst)
        assert np.array_equal(o, o2) and np.array_equal(b, b2.astype(np.uint32))
        assert b2.dtype == (np.uint8 if all(s.isascii() for s in lst) and len(lst) else np.uint32)
    for bad in (["a", None], ["a", 3]):
        with pytest.raises(TypeError):
            strings.pack_strings(bad)


@pytest.mark.parametrize("rng", [(1, 1), (1, 3), (3, 3), (3, 6)])
def test_slot_bounds_cover_every_ngram(rng):
    strs = ["apple", "", "a", "hello  world!!", "x" * 300, "İK"]
    _, offs = pack_utf32(strs)
    slots, occ = ngram_slot_bounds(offs, *rng)
    for s, b in zip(strs, slots):
        for clean in (True, False):
            for rs in (True, False):
                assert len(otfidf.create_ngrams(s, rng, clean, rs)) <= b
    assert occ[-1] == slots.sum() and occ[0] == 0


def test_c_packer_one_pass_paths_equal_the_numpy_paths(monkeypatch):
    """fill_ascii (one pass, guessed buffer) must fall back cleanly -- buffer too small, a non-ASCII string in the middle -- and
    the C slot bounds must equal the numpy expression for every n-gram range."""
    from polyfuzz_b200 import strings, synth
    assert hasattr(strings._hostpack(), "fill_ascii") and hasattr(strings._hostpack(), "slots")
    names = synth.company_names(3000, seed=7)
    ref_b, ref_o = pack_utf32(names)
    for guess in (1.0, 40.0, 4000.0):                       # too small (falls back to scan + fill) / typical / far too large (copies down)
        monkeypatch.setattr(strings, "_BYTES_PER_STRING", guess)
        b, o, _ = strings.pack_strings(names)
        assert b.dtype == np.uint8 and np.array_equal(b.astype(np.uint32), ref_b) and np.array_equal(o, ref_o)
    mixed = names[:100] + ["caf\u00e9 ltd"] + names[100:200]
    b, o, _ = strings.pack_strings(mixed)
    rb, ro = pack_utf32(mixed)
    assert b.dtype == np.uint32 and np.array_equal(b, rb) and np.array_equal(o, ro)
    for rng_ in ((1, 1), (1, 3), (3, 3), (2, 5), (3, 6)):
        s1, o1 = ngram_slot_bounds(ref_o, *rng_)
        monkeypatch.setattr(strings, "_HP", None)            # numpy path
        s2, o2 = ngram_slot_bounds(ref_o, *rng_)
        monkeypatch.undo()
        assert np.array_equal(s1, s2) and np.array_equal(o1, o2)


def test_assemble_matches_equals_reference_tail_restatement():
    rng = np.random.default_rng(0)
    frm = [f"f{i}" for i in range(50)]; to = [f"t{i}" for i in range(20)]
    idx = rng.integers(-1, 20, (50, 3)).astype(np.int32)
    val = np.where(idx >= 0, rng.random((50, 3)), 0.0)
    val[3, 0] = 0.0004
    got = assemble_matches(frm, to, idx, val)
    exp = oracle_assemble(frm, to, idx, val)
    assert list(got.columns) == ["From", "To", "Similarity", "To_2", "Similarity_2", "To_3", "Similarity_3"]
    for c in got.columns:
        g = [None if (isinstance(v, float) and np.isnan(v)) else v for v in got[c].tolist()]
        e = [None if (isinstance(v, float) and np.isnan(v)) else v for v in exp[c].tolist()]
        assert g == e
    assert got["To"][3] is None or got["To"].isna()[3]
    assert clip_top_n(10, ["a", "b", "a"]) == 2 and clip_top_n(10, None) == 10


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    assert lib.pfz_abi_version() == 1
    header = open(os.path.join(ROOT, "include", "pfz.h")).read()
    declared = set(re.findall(r"\b(pfz_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.exported_names()), declared ^ set(_lib.exported_names())
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pfz_scan_ws_bytes(10) >= 256
    assert _lib.launch_count() >= 0
    # every prototype of the header has the same number of parameters as its ctypes binding
    hc = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = re.findall(r"\b(?:int|int64_t|const char \*)\s*\*?\s*(pfz_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", hc, flags=re.S)
    assert {n for n, _ in protos} == declared
    for name, args in protos:
        args = args.strip()
        n_args = 0 if args in ("void", "") else len(args.split(","))
        assert n_args == len(_lib._PROTOS[name]), (name, n_args, len(_lib._PROTOS[name]))


def test_plugin_contract():
    from polyfuzz_b200.matchers import BaseMatcher, TFIDF
    with pytest.raises(TypeError):
        BaseMatcher()                                      # tests/models/test_base.py:29-31
    m = TFIDF(n_gram_range=(2, 3), min_similarity=0.5, top_n=3, model_id="x")
    assert isinstance(m, BaseMatcher) and m.type == "TF-IDF" and m.model_id == "x"
    assert m.vectorizer is None and m.tf_idf_to is None


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="CUDA|cuda"):
        polyfuzz_b200.TFIDF().match(["apple", "apples"], ["apple"])


def test_synth_generators_are_seeded():
    from polyfuzz_b200 import synth
    a = synth.company_names(200, seed=3); b = synth.company_names(200, seed=3)
    assert a == b and a != synth.company_names(200, seed=4)
    u = synth.uniform_strings(100, seed=0)
    assert all(8 <= len(s) <= 32 for s in u)
    t = synth.titles(100, seed=1)
    assert len(t) == 100 and all(len(s) > 0 for s in t)


def test_embeddings_matcher_control_flow_with_stubbed_kernels(monkeypatch):
    """Host logic of Embeddings.match (which vectors are used when, fit/transform state, frame) with the two
    device entry points replaced by numpy stand-ins -- the kernels themselves are covered by tests/test_gpu_dense.py."""
    from polyfuzz_b200 import dense, Embeddings
    calls = []

    def fake_rows(x, normalize=True):
        x = np.asarray(x, dtype=np.float64)
        calls.append(x.shape)
        return x / np.linalg.norm(x, axis=1, keepdims=True), None

    def fake_topk(x, y, k, min_similarity=0.0, self_match=False, **kw):
        s = x @ y.T
        if self_match:
            np.fill_diagonal(s, -np.inf)
        idx = np.argsort(-s, axis=1, kind="stable")[:, :k]
        val = np.take_along_axis(s, idx, 1)
        idx = np.where(val > min_similarity, idx, -1); val = np.where(idx >= 0, val, 0.0)
        return torch.from_numpy(idx.astype(np.int32)), torch.from_numpy(val)

    monkeypatch.setattr(dense, "to_bf16_rows", fake_rows)
    monkeypatch.setattr(dense, "dense_topk", fake_topk)
    rng = np.random.default_rng(0)
    ef, et = rng.normal(size=(4, 8)), rng.normal(size=(3, 8))
    et[1] = ef[2] * 2.0
    frm, to = ["a", "b", "c", "d"], ["x", "y", "z"]
    m = Embeddings(min_similarity=0.0, top_n=2)
    df = m.match(frm, to, ef, et)
    assert list(df.columns) == ["From", "To", "Similarity", "To_2", "Similarity_2"] and df["To"][2] == "y" and df["Similarity"][2] == 1.0
    assert m.embeddings_to is et and calls == [(4, 8), (3, 8)]
    calls.clear()
    df2 = m.match(["q"], to, ef[:1], re_train=False)                 # transform: reuses the fitted to-embeddings
    assert calls == [(1, 8), (3, 8)] and len(df2) == 1
    calls.clear()
    df3 = m.match(frm, None, ef)                                      # self-match: one matrix for both sides
    assert calls == [(4, 8)] and (df3["To"] != df3["From"]).all()
    with pytest.raises(NotImplementedError):
        Embeddings().match(frm, to)                                   # no vectors, no embedder
    emb = Embeddings(embedding_method=lambda strs: rng.normal(size=(len(strs), 8)))
    assert len(emb.match(frm, to)) == 4


def test_arrow_view_of_a_staged_ascii_list():
    """The From column of the device frame tail is a zero-copy Arrow view of the host packer's bytes + offsets."""
    from types import SimpleNamespace
    from polyfuzz_b200.matchers._utils import arrow_from_staged, device_tail_available
    from polyfuzz_b200.strings import pack_strings
    lst = ["alpha", "", "b c", "Zeta-9"]
    blob, off, _ = pack_strings(lst)
    assert blob.dtype == np.uint8
    S = SimpleNamespace(n=len(lst), ascii=True, host_blob=blob, host_off=off)
    assert arrow_from_staged(S).to_pylist() == lst and arrow_from_staged(S, 1, 3).to_pylist() == lst[1:3]
    assert device_tail_available(S) and not device_tail_available(S, None) and not device_tail_available(SimpleNamespace(ascii=False))
