"""GPU parity tests of K1 (vectoriser) and K2 (sparse cosine top-k) against the oracle and against
the golden fixtures generated from the unmodified reference (tests/golden/make_golden.py).
Bit-exact: CSR structure, CSR values, top-k indices and (unrounded) scores."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import native as onative
from oracle import tfidf as otfidf
from oracle.assemble import assemble as oracle_assemble

pytestmark = pytest.mark.gpu

FROM = ["apple", "apples", "appl", "recal", "house", "similarity"]
TO = ["apple", "apples", "mouse"]
RANGES = [(1, 1), (1, 2), (1, 3), (2, 2), (2, 3), (3, 3), (3, 6)]


@pytest.fixture(scope="module")
def pf():
    import polyfuzz_b200
    from polyfuzz_b200 import engine
    return polyfuzz_b200, engine


def _csr_eq(dev_csr, indptr, indices, data, shape):
    m = dev_csr.to_scipy()
    assert m.shape == tuple(shape)
    np.testing.assert_array_equal(m.indptr, indptr)
    np.testing.assert_array_equal(m.indices, indices)
    np.testing.assert_array_equal(m.data, data)          # bit-exact fp64


@pytest.mark.parametrize("lo,hi", RANGES)
@pytest.mark.parametrize("clean", [True, False])
@pytest.mark.parametrize("rs", [True, False])
def test_c1_vectoriser_matches_reference(pf, golden_dir, lo, hi, clean, rs):
    _, engine = pf
    g = np.load(os.path.join(golden_dir, "c1_tfidf.npz"))
    voc = json.load(open(os.path.join(golden_dir, "c1_vocab.json")))
    tag = f"r{lo}{hi}_c{int(clean)}_s{int(rs)}"
    v = engine.NgramTfidf((lo, hi), clean, rs)
    rows_to, rows_from = v.fit_rows([TO, FROM])
    assert v.vocabulary() == voc[tag + "_two"]
    np.testing.assert_array_equal(v.idf, g[tag + "_two_idf"])
    for name, rows in (("to", rows_to), ("from", rows_from)):
        _csr_eq(v.emit(rows), g[f"{tag}_two_{name}_indptr"], g[f"{tag}_two_{name}_indices"], g[f"{tag}_two_{name}_data"],
                g[f"{tag}_two_{name}_shape"])
    v2 = engine.NgramTfidf((lo, hi), clean, rs)
    (rows,) = v2.fit_rows([FROM])
    assert v2.vocabulary() == voc[tag + "_self"]
    _csr_eq(v2.emit(rows), g[tag + "_self_indptr"], g[tag + "_self_indices"], g[tag + "_self_data"], g[tag + "_self_shape"])


def _frame_eq(df, ref):
    assert list(df.columns) == list(ref.keys())
    for c in df.columns:
        got = [None if (isinstance(v, float) and np.isnan(v)) else v for v in df[c].tolist()]
        assert got == ref[c], c


@pytest.mark.parametrize("top_n", [1, 2, 3])
@pytest.mark.parametrize("ms", [0.0, 0.75])
def test_c1_match_frames_equal_reference(pf, golden_dir, top_n, ms):
    polyfuzz_b200, _ = pf
    ref = json.load(open(os.path.join(golden_dir, "c1_match.json")))
    m = polyfuzz_b200.TFIDF(min_similarity=ms, top_n=top_n)
    # the reference sklearn branch ignores min_similarity; compare where they agree: ms=0
    if ms == 0.0:
        _frame_eq(m.match(FROM, TO), ref[f"two_top{top_n}_ms{ms}"])
        m = polyfuzz_b200.TFIDF(min_similarity=ms, top_n=top_n)
        _frame_eq(m.match(FROM), ref[f"self_top{top_n}_ms{ms}"])
    else:
        df = m.match(FROM, TO)
        r0 = ref[f"two_top{top_n}_ms0.0"]
        # sparse-branch semantics: keep only scores > min_similarity
        for c in [c for c in df.columns if c.startswith("Similarity")]:
            exp = [v if v > ms else 0.0 for v in r0[c]]
            assert df[c].tolist() == exp


@pytest.mark.parametrize("method", ["sklearn", "knn"])
@pytest.mark.parametrize("top_n", [1, 2, 3])
def test_c1_sklearn_knn_semantics_ignore_min_similarity(pf, golden_dir, top_n, method):
    """The reference's sklearn / knn branches never apply min_similarity (polyfuzz/models/_utils.py:59-70, 94-102);
    the golden ms=0.75 frames were produced by the unmodified reference through its sklearn branch."""
    polyfuzz_b200, _ = pf
    ref = json.load(open(os.path.join(golden_dir, "c1_match.json")))
    m = polyfuzz_b200.TFIDF(min_similarity=0.75, top_n=top_n, cosine_method=method)
    _frame_eq(m.match(FROM, TO), ref[f"two_top{top_n}_ms0.75"])
    m = polyfuzz_b200.TFIDF(min_similarity=0.75, top_n=top_n, cosine_method=method)
    _frame_eq(m.match(FROM), ref[f"self_top{top_n}_ms0.75"])
    with pytest.raises(ValueError):
        polyfuzz_b200.TFIDF(cosine_method="annoy").match(FROM, TO)


def test_c1_transform_path(pf, golden_dir):
    polyfuzz_b200, _ = pf
    ref = json.load(open(os.path.join(golden_dir, "c1_match.json")))["transform_unseen"]
    m = polyfuzz_b200.TFIDF(min_similarity=0, top_n=1)
    m.match(FROM, TO)
    _frame_eq(m.match(["apples", "mouses", "zzz"], TO, re_train=False), ref)


@pytest.fixture(scope="module")
def company(golden_dir):
    g = np.load(os.path.join(golden_dir, "company_slice.npz"))
    names = json.load(open(os.path.join(golden_dir, "company_slice_names.json")))["names"]
    return g, names


def test_company_slice_vectoriser_bit_exact(pf, company):
    _, engine = pf
    g, names = company
    v = engine.NgramTfidf((3, 3), True, True)
    (rows,) = v.fit_rows([names])
    np.testing.assert_array_equal(v.idf, g["idf"])
    _csr_eq(v.emit(rows), g["csr_indptr"], g["csr_indices"], g["csr_data"], g["csr_shape"])


@pytest.mark.parametrize("variant", ["list", "dense", "dense32", "block"])
@pytest.mark.parametrize("tile", [None, 256, 1024])
@pytest.mark.parametrize("n_splits", [1, 3])
def test_company_slice_topk_vs_oracle_and_reference(pf, company, tile, n_splits, variant):
    polyfuzz_b200, engine = pf
    g, names = company
    k = 10
    v = engine.NgramTfidf((3, 3), True, True)
    (rows,) = v.fit_rows([names])
    csr = v.emit(rows)
    index = engine.SparseIndex(csr, tile=tile, variant=variant)
    idx, val = engine.spcos_topk(csr, index, k, 0.0, self_match=True, n_splits=n_splits)
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    a = csr.to_scipy()
    oi, ov = onative.spdot_topn(a, a, k, 0.0, self_match=True)
    np.testing.assert_array_equal(idx, oi)               # canonical contract: bit-exact indices
    np.testing.assert_array_equal(val, ov)               # and bit-exact fp64 scores
    # unmodified reference (sklearn branch): all 3-dp scores equal; indices equal on tie-free rows
    np.testing.assert_array_equal(np.round(val, 3), g["ref_sims"])
    tf = g["ref_tiefree"]
    nz = g["ref_topvals"] > 0
    same = (idx == g["ref_idx"]) | ~nz
    assert same[tf].all()


def test_company_slice_match_frame(pf, company):
    polyfuzz_b200, engine = pf
    g, names = company
    m = polyfuzz_b200.TFIDF(min_similarity=0, top_n=10)
    df = m.match(names)
    a = m.tf_idf_to.to_scipy()
    oi, ov = onative.spdot_topn(a, a, 10, 0.0, self_match=True)
    exp = oracle_assemble(names, None, oi, ov)
    assert list(df.columns) == list(exp.columns)
    for c in df.columns:
        if c.startswith("Similarity"):
            np.testing.assert_array_equal(df[c].to_numpy(), exp[c].to_numpy())
        else:
            assert [None if (isinstance(v, float) and np.isnan(v)) else v for v in df[c].tolist()] == \
                   [None if (isinstance(v, float) and np.isnan(v)) else v for v in exp[c].tolist()]


def _oracle_two(frm, to, rng=(3, 3), clean=True, rs=True):
    f, t, _ = otfidf.fit_transform_sklearn(frm, to, rng, clean, rs)
    return f, t


def _force_variant(monkeypatch, engine, variant):
    monkeypatch.setattr(engine, "DENSE_MIN_DENSITY", 1e9 if variant == "list" else 0.0)
    monkeypatch.setattr(engine, "DENSE_VARIANT", variant if variant != "list" else "dense")
    if variant != "block":                                             # (the block kernel's tables hold <= 128 terms per row)
        monkeypatch.setattr(engine, "DENSE32_MAX_ROW_NNZ", 1 << 30)    # exercise the filter even on long rows


@pytest.mark.parametrize("variant", ["list", "dense", "dense32", "block"])
@pytest.mark.parametrize("k,ms", [(1, 0.0), (5, 0.3), (32, 0.0), (40, 0.0), (70, 0.05)])
def test_synthetic_two_list_vs_oracle(pf, k, ms, variant, monkeypatch):
    polyfuzz_b200, engine = pf
    from polyfuzz_b200 import synth
    _force_variant(monkeypatch, engine, variant)
    to = synth.company_names(6000, seed=3)
    frm = synth.company_names(2500, seed=4) + ["", "a", "ab", "  ", "!!!", to[17], to[17].lower()]
    m = polyfuzz_b200.TFIDF(min_similarity=ms, top_n=k)
    idx, val, kk = m.match_arrays(frm, to)
    f, t = _oracle_two(frm, to)
    np.testing.assert_array_equal(m.tf_idf_to.to_scipy().data, t.data)
    oi, ov = onative.spdot_topn(f, t, kk, ms)
    np.testing.assert_array_equal(idx.cpu().numpy(), oi)
    np.testing.assert_array_equal(val.cpu().numpy(), ov)


def test_uniform_strings_vs_oracle(pf):
    """BASELINE config 5 shape (uniform 8..32 chars over 37 symbols) at an oracle-sized scale."""
    polyfuzz_b200, engine = pf
    from polyfuzz_b200 import synth
    to = synth.uniform_strings(20000, seed=0)
    frm = synth.uniform_strings(5000, seed=1)
    m = polyfuzz_b200.TFIDF(min_similarity=0.0, top_n=10)
    idx, val, kk = m.match_arrays(frm, to)
    f, t = _oracle_two(frm, to)
    oi, ov = onative.spdot_topn(f, t, kk, 0.0, n_threads=8)
    np.testing.assert_array_equal(idx.cpu().numpy(), oi)
    np.testing.assert_array_equal(val.cpu().numpy(), ov)


def test_edge_cases(pf):
    polyfuzz_b200, engine = pf
    # long strings (> 256 n-gram slots -> long-row kernel), duplicates, unicode survivors, whitespace
    long_a = "alpha beta gamma delta " * 30
    long_b = "alpha beta gamma delta " * 29 + "epsilon"
    frm = [long_a, "İstanbul Kelvin K", "x\ty\nz  w", "dup", "dup", "", "ab"]
    to = [long_b, "istanbul kelvin k", "xyz w", "dup", "other", "dup"]
    for rng in [(3, 3), (2, 4), (1, 1)]:
        for clean in (True, False):
            m = polyfuzz_b200.TFIDF(n_gram_range=rng, clean_string=clean, min_similarity=0.0, top_n=3)
            idx, val, kk = m.match_arrays(frm, to)
            f, t = _oracle_two(frm, to, rng, clean, True)
            got = m.tf_idf_to.to_scipy()
            np.testing.assert_array_equal(got.indices, t.indices)
            np.testing.assert_array_equal(got.data, t.data)
            oi, ov = onative.spdot_topn(f, t, kk, 0.0)
            np.testing.assert_array_equal(idx.cpu().numpy(), oi)
            np.testing.assert_array_equal(val.cpu().numpy(), ov)
    # self-match excludes only the diagonal: duplicates at other positions still match with 1.0
    m = polyfuzz_b200.TFIDF(min_similarity=0.0, top_n=1)
    df = m.match(["dup", "dup", "zzz"])
    assert df["To"].tolist()[:2] == ["dup", "dup"] and df["Similarity"].tolist()[:2] == [1.0, 1.0]
    assert df["To"].tolist()[2] is None or df["To"].isna().tolist()[2]
    # top_n is clipped to the number of distinct to-strings (_utils.py:54-56)
    m = polyfuzz_b200.TFIDF(min_similarity=0.0, top_n=10)
    assert list(m.match(FROM, TO).columns) == ["From", "To", "Similarity", "To_2", "Similarity_2", "To_3", "Similarity_3"]
    # empty vocabulary raises like scikit-learn
    with pytest.raises(ValueError, match="empty vocabulary"):
        polyfuzz_b200.TFIDF().match(["a", "b"], ["c"])


def test_pickle_round_trip(pf, tmp_path):
    import joblib
    polyfuzz_b200, _ = pf
    m = polyfuzz_b200.TFIDF(min_similarity=0, top_n=1)
    m.match(FROM, TO)
    exp = m.match(["appl", "mouses"], TO, re_train=False)
    joblib.dump(m, tmp_path / "m.joblib")
    m2 = joblib.load(tmp_path / "m.joblib")
    got = m2.match(["appl", "mouses"], TO, re_train=False)
    assert got.equals(exp)


def test_from_block_is_a_row_block_of_the_self_match(pf):
    polyfuzz_b200, engine = pf
    from polyfuzz_b200 import synth
    names = synth.company_names(3000, seed=9)
    full = polyfuzz_b200.TFIDF(min_similarity=0.0, top_n=4).match(names)
    blk = polyfuzz_b200.TFIDF(min_similarity=0.0, top_n=4).match(names, from_block=(1000, 1800))
    assert blk.reset_index(drop=True).equals(full.iloc[1000:1800].reset_index(drop=True))


@pytest.mark.parametrize("variant", ["list", "dense", "dense32", "block"])
def test_every_row_shares_many_heavy_terms(pf, variant, monkeypatch):
    """All rows share ~20 trigrams, so every (term, tile) segment is the full tile: per-unit work far
    exceeds the dense kernel's work-item table (batched consumption) and every accumulator gets ~20
    additions in term order."""
    polyfuzz_b200, engine = pf
    _force_variant(monkeypatch, engine, variant)
    names = [f"alpha beta gamma delta {i:05d} {'x' * (i % 7)}{i % 13}" for i in range(2600)]
    m = polyfuzz_b200.TFIDF(min_similarity=0.0, top_n=7)
    idx, val, k = m.match_arrays(names)
    a = m.tf_idf_to.to_scipy()
    oi, ov = onative.spdot_topn(a, a, k, 0.0, self_match=True, n_threads=8)
    np.testing.assert_array_equal(idx.cpu().numpy(), oi)
    np.testing.assert_array_equal(val.cpu().numpy(), ov)


def test_shard_emulation_equals_unsharded(pf):
    """Multi-GPU result = merge of per-shard top-k with global indices (emulated sequentially on one GPU:
    same vectoriser state on every 'rank', one index per to-block, pfz_topk_merge)."""
    import torch
    polyfuzz_b200, engine = pf
    from polyfuzz_b200 import synth
    from polyfuzz_b200.distributed import shard_bounds
    to = synth.company_names(5000, seed=21); frm = synth.company_names(1200, seed=22)
    v = engine.NgramTfidf((3, 3), True, True)
    rows_to, rows_from = v.fit_rows([to, frm])
    csr_to, csr_from = v.emit(rows_to), v.emit(rows_from)
    full_i, full_v = engine.spcos_topk(csr_from, engine.SparseIndex(csr_to, variant="dense"), 10, 0.0)
    parts_i, parts_v = [], []
    for r in range(3):
        lo, hi = shard_bounds(len(to), 3, r)
        blk = v.transform(to[lo:hi])
        i_, v_ = engine.spcos_topk(csr_from, engine.SparseIndex(blk, variant="dense"), 10, 0.0, to_index_base=lo)
        parts_i.append(i_); parts_v.append(v_)
    mi, mv = engine.topk_merge(torch.stack(parts_i), torch.stack(parts_v), 10)
    assert torch.equal(mi, full_i) and torch.equal(mv, full_v)


@pytest.mark.parametrize("variant", ["dense32", "block", "block16", "block32", "block16x4"])
def test_dense32_many_exact_ties_and_identical_rows(pf, monkeypatch, variant):
    """The fp32 filter must hand every row near the k-th key to the exact re-scoring: lists full of duplicates
    (scores exactly 1.0 and large groups of exactly tied scores)."""
    polyfuzz_b200, engine = pf
    if variant.startswith("block") and variant != "block":
        monkeypatch.setattr(engine, "BLOCK_ACC_BITS", 32 if variant == "block32" else 16)
        monkeypatch.setattr(engine, "BLOCK_ROWS", 4 if variant.endswith("x4") else 8); variant = "block"
    _force_variant(monkeypatch, engine, variant)
    names = ["acme holdings inc"] * 40 + ["dup dup llc"] * 600 + ["acme holding inc"] * 40 + [f"zeta {i % 5} llc" for i in range(300)] + ["unique name ltd"]
    m = polyfuzz_b200.TFIDF(min_similarity=0.0, top_n=12)
    idx, val, k = m.match_arrays(names)
    a = m.tf_idf_to.to_scipy()
    oi, ov = onative.spdot_topn(a, a, k, 0.0, self_match=True)
    np.testing.assert_array_equal(idx.cpu().numpy(), oi)
    np.testing.assert_array_equal(val.cpu().numpy(), ov)


def test_variant_selection_rules(pf):
    _, engine = pf
    assert engine.choose_variant(0.001, 20) == "list" and engine.choose_variant(None, 20) == "list"
    assert engine.choose_variant(0.001, 20, 1000) == "list" and engine.choose_variant(0.001, 20, 1_000_000) == "hash"
    assert engine.choose_variant(0.001, 2000, 1_000_000) == "list"
    assert engine.choose_variant(0.3, 69) == engine.DENSE_VARIANT
    assert engine.choose_variant(0.3, 5000) == "dense" and engine.choose_variant(0.3, None) == "dense"


@pytest.mark.parametrize("rng,clean", [((3, 6), True), ((2, 5), True), ((1, 3), False), ((3, 4), False)])
def test_large_code_space_and_raw_mode_at_scale(pf, rng, clean):
    """n-gram ranges whose code space exceeds the direct-addressed table (gather + bitonic sort + RLE vocabulary)
    and the raw (clean_string=False) alphabet path, on a few thousand strings with non-ASCII letters."""
    polyfuzz_b200, engine = pf
    from polyfuzz_b200 import synth
    to = synth.titles(2500, seed=31); frm = synth.titles(700, seed=32) + ["", "x", "Ünïcødé Çafé", to[3]]
    m = polyfuzz_b200.TFIDF(n_gram_range=rng, clean_string=clean, min_similarity=0.0, top_n=5)
    idx, val, k = m.match_arrays(frm, to)
    f, t = _oracle_two(frm, to, rng, clean, True)
    got_to = m.tf_idf_to.to_scipy()
    assert got_to.shape == t.shape
    np.testing.assert_array_equal(got_to.indptr, t.indptr)
    np.testing.assert_array_equal(got_to.indices, t.indices)
    np.testing.assert_array_equal(got_to.data, t.data)
    oi, ov = onative.spdot_topn(f, t, k, 0.0, n_threads=8)
    np.testing.assert_array_equal(idx.cpu().numpy(), oi)
    np.testing.assert_array_equal(val.cpu().numpy(), ov)


def test_transform_after_pickle_with_unseen_and_empty_rows(pf, tmp_path):
    import joblib
    polyfuzz_b200, engine = pf
    from polyfuzz_b200 import synth
    to = synth.company_names(3000, seed=41); frm = synth.company_names(500, seed=42)
    m = polyfuzz_b200.TFIDF(min_similarity=0.0, top_n=3)
    m.match(frm, to)
    new = synth.company_names(300, seed=43) + ["", "zzzzqqqq xxxyyy", "A"]
    exp = m.match(new, to, re_train=False)
    # oracle: fixed vocabulary / idf from the fit, unseen n-grams dropped
    o = otfidf.TfidfOracle().fit(list(to) + list(frm))
    oi, ov = onative.spdot_topn(o.transform(new), o.transform(to), 3, 0.0)
    ref = oracle_assemble(new, to, oi, ov)
    for c in exp.columns:
        g = [None if (isinstance(v, float) and np.isnan(v)) else v for v in exp[c].tolist()]
        e = [None if (isinstance(v, float) and np.isnan(v)) else v for v in ref[c].tolist()]
        assert g == e, c
    joblib.dump(m, tmp_path / "m.joblib")
    got = joblib.load(tmp_path / "m.joblib").match(new, to, re_train=False)
    assert got.equals(exp)


@pytest.mark.parametrize("tag,rng,clean", [("raw33", (3, 3), False), ("clean13", (1, 3), True), ("raw12", (1, 2), False)])
def test_titles_slice_vectoriser_matches_reference_on_gpu(pf, golden_dir, tag, rng, clean):
    """Real movie titles (non-ASCII letters, punctuation, digits; raw and clean mode) from the reference's
    data/movie_titles.json: K1 reproduces the unmodified reference's CSR and idf bit for bit (the CPU twin of this test,
    tests/test_oracle_golden.py, pins the oracle on the same fixture)."""
    _, engine = pf
    g = np.load(os.path.join(golden_dir, "titles_slice.npz"))
    names = json.load(open(os.path.join(golden_dir, "titles_slice_names.json")))
    frm, to = names["from"], names["to"]
    v = engine.NgramTfidf(rng, clean, True)
    rows_to, rows_from = v.fit_rows([to, frm])
    np.testing.assert_array_equal(v.idf, g[tag + "_idf"])
    for name, rows in (("from", rows_from), ("to", rows_to)):
        _csr_eq(v.emit(rows), g[f"{tag}_{name}_indptr"], g[f"{tag}_{name}_indices"], g[f"{tag}_{name}_data"], g[f"{tag}_{name}_shape"])


@pytest.mark.parametrize("acc_bits,rows", [(32, 8), (16, 8), (16, 16), (32, 16), (16, 4), (32, 4)])
def test_block_variant_long_rows_split_blocks_and_margin(pf, monkeypatch, acc_bits, rows):
    """The from-row-block kernel at its contract's edge: rows of ~100-128 distinct trigrams (blocks of 8 such rows exceed the
    512-entry block table and are split in halves), mixed with short rows, duplicates and empty rows; tile 128 and a tile
    split.  The fp32 filter (margin 3e-5) must still hand every contender to the exact re-scoring."""
    polyfuzz_b200, engine = pf
    from polyfuzz_b200 import synth
    monkeypatch.setattr(engine, "BLOCK_ACC_BITS", acc_bits)
    monkeypatch.setattr(engine, "BLOCK_ROWS", rows)
    rng = np.random.default_rng(5)
    words = synth.company_names(400, seed=12)
    long_rows = [" ".join(rng.choice(words, 5)) for _ in range(300)]
    names = synth.company_names(2000, seed=13) + long_rows + ["", "ab", long_rows[0], long_rows[0] + " x"]
    v = engine.NgramTfidf((3, 3), True, True)
    (rows,) = v.fit_rows([names]); csr = v.emit(rows)
    a = csr.to_scipy()
    nnz = np.diff(a.indptr)
    assert 100 < nnz.max() <= 128, nnz.max()
    oi, ov = onative.spdot_topn(a, a, 10, 0.0, self_match=True, n_threads=8)
    for tile, splits in ((128, 1), (1024, 1), (512, 3), (256, 1), (2048, 1)):
        ix = engine.SparseIndex(csr, tile=tile, variant="block")
        idx, val = engine.spcos_topk(csr, ix, 10, 0.0, self_match=True, n_splits=splits)
        assert int(ix._block_err.item()) == 0
        np.testing.assert_array_equal(idx.cpu().numpy(), oi)
        np.testing.assert_array_equal(val.cpu().numpy(), ov)


@pytest.mark.parametrize("slots", [2048, 8192])
@pytest.mark.parametrize("k,ms,self_match", [(10, 0.0, False), (3, 0.05, False), (40, 0.0, False), (10, 0.0, True)])
def test_hash_variant_uniform_strings_vs_oracle(pf, monkeypatch, slots, k, ms, self_match):
    """The sparse-regime hash kernel (BASELINE config 5 shape at an oracle-sized scale): several 65 536-row tiles, table
    sizes that force multi-pass rows, paging (top_n > 32), a threshold, and the self-match diagonal."""
    polyfuzz_b200, engine = pf
    from polyfuzz_b200 import synth
    monkeypatch.setattr(engine, "HASH_SLOTS", slots)
    to = synth.uniform_strings(150_000, seed=0)
    frm = to[:3000] if self_match else synth.uniform_strings(3000, seed=1) + ["", "zz", to[5], to[70000]]
    v = engine.NgramTfidf((3, 3), True, True)
    rows_to, rows_from = v.fit_rows([to, frm])
    csr_to, csr_from = v.emit(rows_to), v.emit(rows_from)
    ix = engine.SparseIndex(csr_to, variant="hash")
    assert ix.tile == 65536 and ix.n_tiles == 3
    idx, val = engine.spcos_topk(csr_from, ix, k, ms, self_match=self_match)
    assert int(ix._hash_err.item()) == 0
    oi, ov = onative.spdot_topn(csr_from.to_scipy(), csr_to.to_scipy(), k, ms, self_match=self_match, n_threads=16)
    np.testing.assert_array_equal(idx.cpu().numpy(), oi)
    np.testing.assert_array_equal(val.cpu().numpy(), ov)


def test_hash_variant_dense_rows_never_overflow_the_table(pf, monkeypatch):
    """The hash kernel on input far denser than its cost model assumes (company names: a from-row visits ~20 000 postings per
    65 536-row tile against a 2 048-slot table): passes whose table fills are redone over halved to-row ranges, results stay
    bit-identical to the oracle and no error flag is raised."""
    polyfuzz_b200, engine = pf
    from polyfuzz_b200 import synth
    monkeypatch.setattr(engine, "HASH_SLOTS", 2048)
    to = synth.company_names(70_000, seed=3)
    frm = synth.company_names(200, seed=4) + [to[5], to[69_999], ""]
    v = engine.NgramTfidf((3, 3), True, True)
    rows_to, rows_from = v.fit_rows([to, frm])
    csr_to, csr_from = v.emit(rows_to), v.emit(rows_from)
    ix = engine.SparseIndex(csr_to, variant="hash")
    assert ix.n_tiles == 2
    idx, val = engine.spcos_topk(csr_from, ix, 10, 0.0)
    assert int(ix._hash_err.item()) == 0
    oi, ov = onative.spdot_topn(csr_from.to_scipy(), csr_to.to_scipy(), 10, 0.0, self_match=False, n_threads=16)
    np.testing.assert_array_equal(idx.cpu().numpy(), oi)
    np.testing.assert_array_equal(val.cpu().numpy(), ov)
