"""Parity at BASELINE.json's full sizes ON THE REFERENCE'S OWN INPUTS (data/company_names.json and
data/movie_titles.json travel as fixtures under tests/golden/data/), through properties that do not need the oracle on the whole grid plus the
oracle on a sample of rows:
  config 2 (100k x 100k TF-IDF top-10): the three K2 kernels agree bit for bit; lists are sorted, duplicate-free and
      diagonal-free; scores are symmetric bit for bit (score(i,j) == score(j,i): same products, same order) and a row's
      list contains every reverse neighbour that beats its k-th key; 512 sampled rows equal the oracle exactly;
  config 3 (6172 x 80852 edit distance): sampled rows equal the DP oracle; Levenshtein symmetry on a sub-grid;
  config 4 (100k x 100k x 768 dense): sampled rows vs fp64 on the same bf16 inputs within 1e-5."""
import numpy as np
import pytest
import torch

from oracle import native as onative
from oracle import tfidf as otfidf

pytestmark = pytest.mark.gpu


def test_config2_full_size_properties():
    import polyfuzz_b200
    from polyfuzz_b200 import datasets, engine
    n, k = 100_000, 10
    names, kind = datasets.load_company_names(n)          # the reference's data/company_names.json (fixture)
    assert kind == "real" and len(names) == n
    v = engine.NgramTfidf((3, 3), True, True)
    (rows,) = v.fit_rows([names]); csr = v.emit(rows)
    res = {}
    for variant, tile in (("block", None), ("dense32", None), ("dense", None), ("list", None)):
        ix = engine.SparseIndex(csr, tile=tile, variant=variant)
        res[variant] = engine.spcos_topk(csr, ix, k, 0.0, self_match=True)
    for variant in ("block", "dense", "list"):
        assert torch.equal(res[variant][0], res["dense32"][0]) and torch.equal(res[variant][1], res["dense32"][1]), variant
    idx = res["dense32"][0].cpu().numpy(); val = res["dense32"][1].cpu().numpy()
    valid = idx >= 0
    assert (np.diff(val, axis=1) <= 0).all()
    assert (idx != np.arange(n)[:, None]).all()
    srt = np.sort(np.where(valid, idx, -np.arange(1, k + 1)[None, :]), axis=1)
    assert (np.diff(srt, axis=1) != 0).all()                                  # no duplicates inside a row
    # symmetry: for every (i -> j, s) either j lists i with exactly s, or s does not beat j's k-th key
    ii = np.repeat(np.arange(n), k)[valid.ravel()]; jj = idx.ravel()[valid.ravel()]; ss = val.ravel()[valid.ravel()]
    back = idx[jj] == ii[:, None]
    has = back.any(axis=1)
    assert np.array_equal(val[jj][back], ss[has])                            # bit-identical reverse score
    kth_v = val[jj, k - 1]; kth_i = idx[jj, k - 1]
    beats = (ss > kth_v) | ((ss == kth_v) & (ii < kth_i))
    assert not (beats & ~has).any()
    # oracle on a sample of rows against the whole list
    a = csr.to_scipy()
    sel = np.random.default_rng(0).choice(n, 512, replace=False); sel.sort()
    inv = onative.InvertedIndex(a)
    oi = np.empty((len(sel), k), np.int32); ov = np.empty((len(sel), k))
    for t, r in enumerate(sel):
        o_i, o_v = onative.spdot_topn(a[r:r + 1], inv, k, 0.0, self_match=True, from_index_base=int(r))
        oi[t], ov[t] = o_i[0], o_v[0]
    np.testing.assert_array_equal(idx[sel], oi)
    np.testing.assert_array_equal(val[sel], ov)


def test_config3_full_size_sample_and_symmetry():
    from polyfuzz_b200 import datasets, editdist
    titles, kind = datasets.load_movie_titles()           # the reference's data/movie_titles.json (fixture)
    assert kind == "real"
    frm, to = titles["Netflix"], titles["IMDB"]
    assert (len(frm), len(to)) == (6172, 80852)
    bi, bs, bd = editdist.edit_argbest(frm, to, "norm_lev")
    bi, bs, bd = bi.cpu().numpy(), bs.cpu().numpy(), bd.cpu().numpy()
    sel = np.random.default_rng(1).choice(len(frm), 48, replace=False)
    oi, os_, od = onative.editdist_argbest([frm[i] for i in sel], to, "norm_lev", n_threads=16)
    np.testing.assert_array_equal(bi[sel], oi); np.testing.assert_array_equal(bd[sel], od); np.testing.assert_array_equal(bs[sel], os_)
    assert (bs >= 0).all() and (bs <= 1).all() and (bi >= 0).all()
    a, b = frm[:700], to[:900]
    _, _, _, m1 = editdist.edit_argbest(a, b, "lev", want_matrix=True)
    _, _, _, m2 = editdist.edit_argbest(b, a, "lev", want_matrix=True)
    assert torch.equal(m1, m2.T.contiguous())                                  # d(a,b) == d(b,a)
    la = np.array([len(s) for s in a])[:, None]; lb = np.array([len(s) for s in b])[None, :]
    m = m1.cpu().numpy()
    assert (m >= np.abs(la - lb)).all() and (m <= np.maximum(la, lb)).all()


def test_config4_full_size_sample():
    from polyfuzz_b200 import dense
    n, d, k = 100_000, 768, 10
    dev = torch.device("cuda")
    torch.manual_seed(0); X = torch.randn(n, d, device=dev); torch.manual_seed(1); Y = torch.randn(n, d, device=dev)
    x, _ = dense.to_bf16_rows(X, True); y, _ = dense.to_bf16_rows(Y, True)
    idx, val = dense.dense_topk(x, y, k, 0.0)
    sel = torch.from_numpy(np.random.default_rng(2).choice(n, 256, replace=False)).to(dev)
    s = x[sel].double() @ y.double().T
    rv, ri = torch.topk(s, k, dim=1)
    got_v = val[sel]; got_i = idx[sel].long()
    assert torch.allclose(got_v, rv, atol=1e-5, rtol=0)
    assert torch.allclose(torch.gather(s, 1, got_i), got_v, atol=1e-5, rtol=0)
    same = (got_i == ri).float().mean().item()
    assert same > 0.99                                                         # differences only at fp32/fp64 near-ties
