"""GPU parity tests of K3 (bit-parallel Levenshtein / Indel + arg-best) against the Wagner-Fischer oracle.
Integer distances and arg-best indices are bit-exact; scores are the same IEEE expressions."""
import numpy as np
import pytest

from oracle import native as onative

pytestmark = pytest.mark.gpu

FROM = ["apple", "apples", "appl", "recal", "house", "similarity"]
TO = ["apple", "apples", "mouse"]


@pytest.fixture(scope="module")
def ed():
    from polyfuzz_b200 import editdist
    return editdist


def _rand_strings(rng, n, lo, hi, alpha):
    return ["".join(alpha[i] for i in rng.integers(0, len(alpha), rng.integers(lo, hi + 1))) for _ in range(n)]


@pytest.mark.parametrize("metric", ["lev", "indel"])
def test_c1_matrix_known_answers(ed, metric):
    _, _, _, mat = ed.edit_argbest(FROM, TO, metric, want_matrix=True)
    exp = onative.editdist_matrix(FROM, TO, metric)
    np.testing.assert_array_equal(mat.cpu().numpy(), exp)
    if metric == "lev":
        assert mat.cpu().numpy().tolist() == [[0, 1, 4], [1, 0, 5], [1, 2, 5], [5, 6, 5], [4, 5, 1], [9, 9, 9]]


@pytest.mark.parametrize("metric", ["lev", "indel", "norm_lev", "ratio"])
@pytest.mark.parametrize("lo,hi", [(0, 12), (20, 40), (50, 70), (90, 140), (200, 300), (500, 600)])
def test_random_strings_all_word_classes(ed, metric, lo, hi):
    rng = np.random.default_rng(lo * 7 + hi)
    alpha = "abcdefgh éß中K"
    frm = _rand_strings(rng, 70, lo, hi, alpha) + ["", "a"]
    to = _rand_strings(rng, 150, max(0, lo // 2), hi + 10, alpha) + ["", frm[3], frm[3][:-1] if frm[3] else "x"]
    bi, bs, bd, mat = ed.edit_argbest(frm, to, metric, want_matrix=True, n_splits=3)
    dmetric = "lev" if metric in ("lev", "norm_lev") else "indel"
    np.testing.assert_array_equal(mat.cpu().numpy(), onative.editdist_matrix(frm, to, dmetric, n_threads=8))
    oi, os_, od = onative.editdist_argbest(frm, to, metric, score_cutoff=float("-inf") if metric in ("lev", "indel") else 0.0, n_threads=8)
    np.testing.assert_array_equal(bi.cpu().numpy(), oi)
    np.testing.assert_array_equal(bd.cpu().numpy(), od)
    np.testing.assert_array_equal(bs.cpu().numpy(), os_)


def test_titles_grid_vs_oracle(ed):
    """BASELINE config 3 shape (movie-title-like strings, a few non-ASCII) at an oracle-sized scale."""
    from polyfuzz_b200 import synth
    frm = synth.titles(300, seed=1); to = synth.titles(5000, seed=2)
    for metric in ("ratio", "norm_lev"):
        bi, bs, bd = ed.edit_argbest(frm, to, metric)
        oi, os_, od = onative.editdist_argbest(frm, to, metric, n_threads=8)
        np.testing.assert_array_equal(bi.cpu().numpy(), oi)
        np.testing.assert_array_equal(bd.cpu().numpy(), od)
        np.testing.assert_array_equal(bs.cpu().numpy(), os_)


def test_self_match_cutoff_and_big_alphabet(ed):
    rng = np.random.default_rng(3)
    s = _rand_strings(rng, 400, 3, 20, "abcdef") + ["dup", "dup"]
    bi, bs, bd = ed.edit_argbest(s, s, "ratio", score_cutoff=60.0, exclude_self=True)
    oi, os_, od = onative.editdist_argbest(s, s, "ratio", score_cutoff=60.0, exclude_self=True)
    np.testing.assert_array_equal(bi.cpu().numpy(), oi); np.testing.assert_array_equal(bs.cpu().numpy(), os_)
    assert (bi.cpu().numpy() != np.arange(len(s))).all()
    # more than 255 distinct code points in the from-list -> alphabet batches
    big = [chr(0x4E00 + i) + chr(0x4E00 + (i * 7) % 600) + "ab" for i in range(600)]
    to = big[::3] + ["ab", "中ab"]
    bi, bs, bd = ed.edit_argbest(big, to, "norm_lev")
    oi, os_, od = onative.editdist_argbest(big, to, "norm_lev")
    np.testing.assert_array_equal(bi.cpu().numpy(), oi); np.testing.assert_array_equal(bd.cpu().numpy(), od)
    with pytest.raises(ValueError, match="at most"):
        ed.edit_argbest(["x" * 1025], ["y"], "lev")


def test_matchers_mirror_reference_tests():
    """tests/models/test_rapidfuzz.py and tests/models/test_distance.py of the reference, with fuzz.ratio."""
    import pandas as pd
    from polyfuzz_b200 import RapidFuzz, EditDistance
    m = RapidFuzz(scorer="ratio").match(FROM, TO)
    assert isinstance(m, pd.DataFrame) and len(m) == 6 and list(m.columns) == ["From", "To", "Similarity"]
    assert m.Similarity.mean() > 0.0
    np.testing.assert_allclose(m.Similarity.to_numpy(), np.array([100, 100, 88.88888888888889, 40, 80, 13.33333333333333]) / 100, rtol=1e-12)
    assert m.To.tolist() == ["apple", "apples", "apple", "apple", "mouse", "apple"]
    m = RapidFuzz(score_cutoff=0.95).match(FROM, TO)
    assert m.Similarity.mean() < 0.5 and m.To.tolist()[2:] == [None] * 4 and m.Similarity.tolist() == [1.0, 1.0, 0, 0, 0, 0]
    e = EditDistance(normalize=False).match(FROM, TO)
    assert e.Similarity.mean() > 50 and len(e) == 6 and list(e.columns) == ["From", "To", "Similarity"]
    e = EditDistance().match(FROM, TO)
    assert e.Similarity.min() == 0.0 and e.Similarity.max() == 1.0
    s = RapidFuzz().match(["dup", "dup", "other"])
    assert s.To.tolist()[:2] == ["dup", "dup"] and s.Similarity.tolist()[:2] == [1.0, 1.0]
    with pytest.raises(NotImplementedError):
        RapidFuzz(scorer=lambda a, b: 1.0)
