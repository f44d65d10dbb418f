"""ctypes bindings to oracle/liboracle.so (TEST INFRASTRUCTURE, NOT PRODUCT CODE)."""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

METRIC = {"lev": 0, "indel": 1, "norm_lev": 2, "ratio": 3}


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("spdot_topn.c", "editdist.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class InvertedIndex:
    """to-matrix transposed to term-major posting lists (what awesome_cossim_topn receives as B)."""

    def __init__(self, b_csr):
        import scipy.sparse as sp
        bt = sp.csr_matrix(b_csr).T.tocsr(); bt.sort_indices()       # docs ascending inside each term
        self.n_to = b_csr.shape[0]
        self.indptr = np.ascontiguousarray(bt.indptr, dtype=np.int32)
        self.indices = np.ascontiguousarray(bt.indices, dtype=np.int32)
        self.data = np.ascontiguousarray(bt.data, dtype=np.float64)


def spdot_topn(a_csr, b_csr, k, lower_bound=0.0, self_match=False, from_index_base=0,
               to_index_base=0, n_threads=1):
    """Gustavson sparse dot + canonical top-k (restates awesome_cossim_topn at _utils.py:82 plus
    the post-processing at _utils.py:84-91,128-146).  a_csr: from (n_from x V); b_csr: to (n_to x V)
    or a prebuilt InvertedIndex."""
    import scipy.sparse as sp
    a = sp.csr_matrix(a_csr); a.sort_indices()
    bt = b_csr if isinstance(b_csr, InvertedIndex) else InvertedIndex(b_csr)
    n_from, n_to = a.shape[0], bt.n_to
    ai = np.ascontiguousarray(a.indptr, dtype=np.int32); aj = np.ascontiguousarray(a.indices, dtype=np.int32)
    av = np.ascontiguousarray(a.data, dtype=np.float64)
    idx = np.empty((n_from, k), dtype=np.int32); val = np.empty((n_from, k), dtype=np.float64)
    rc = lib().oracle_spdot_topn(ctypes.c_int32(n_from), ctypes.c_int32(n_to), _p(ai), _p(aj), _p(av),
                                 _p(bt.indptr), _p(bt.indices), _p(bt.data), ctypes.c_int32(k), ctypes.c_double(lower_bound),
                                 ctypes.c_int32(int(self_match)), ctypes.c_int64(from_index_base),
                                 ctypes.c_int64(to_index_base), _p(idx), _p(val), ctypes.c_int32(n_threads))
    if rc:
        raise MemoryError("oracle_spdot_topn")
    return idx, val


def topk_merge(idx, val, k_out):
    idx = np.ascontiguousarray(idx, dtype=np.int32); val = np.ascontiguousarray(val, dtype=np.float64)
    n_shards, n_from, k_in = idx.shape
    oi = np.empty((n_from, k_out), dtype=np.int32); ov = np.empty((n_from, k_out), dtype=np.float64)
    lib().oracle_topk_merge(ctypes.c_int32(n_shards), ctypes.c_int32(n_from), ctypes.c_int32(k_in),
                            ctypes.c_int32(k_out), _p(idx), _p(val), _p(oi), _p(ov))
    return oi, ov


def pack_utf32(strings):
    """list[str] -> (uint32 code points, int64 offsets[n+1]); the layout both the oracle and the
    C-ABI use for string lists."""
    n = len(strings)
    offs = np.zeros(n + 1, dtype=np.int64)
    if n:
        np.cumsum(np.fromiter(map(len, strings), dtype=np.int64, count=n), out=offs[1:])
    blob = np.frombuffer("".join(strings).encode("utf-32-le", "surrogatepass"), dtype=np.uint32)
    if blob.size == 0:
        blob = np.zeros(1, dtype=np.uint32)
    return np.ascontiguousarray(blob), offs


def editdist_matrix(from_list, to_list, metric="lev", n_threads=1):
    fb, fo = pack_utf32(from_list); tb, to = pack_utf32(to_list)
    d = np.empty((len(from_list), len(to_list)), dtype=np.int32)
    lib().oracle_editdist_matrix(_p(fb), _p(fo), ctypes.c_int32(len(from_list)), _p(tb), _p(to),
                                 ctypes.c_int32(len(to_list)), ctypes.c_int32(METRIC[metric]), _p(d),
                                 ctypes.c_int32(n_threads))
    return d


def editdist_argbest(from_list, to_list, metric="ratio", score_cutoff=0.0, exclude_self=False,
                     self_shift=0, n_threads=1, myers=False):
    fb, fo = pack_utf32(from_list); tb, to = pack_utf32(to_list)
    n = len(from_list)
    bi = np.empty(n, dtype=np.int32); bs = np.empty(n, dtype=np.float64); bd = np.empty(n, dtype=np.int32)
    if myers:
        assert metric == "norm_lev" and not exclude_self and score_cutoff <= 0.0
        lib().oracle_lev_argbest_myers(_p(fb), _p(fo), ctypes.c_int32(n), _p(tb), _p(to),
                                       ctypes.c_int32(len(to_list)), _p(bi), _p(bs), _p(bd),
                                       ctypes.c_int32(n_threads))
    else:
        lib().oracle_editdist_argbest(_p(fb), _p(fo), ctypes.c_int32(n), _p(tb), _p(to),
                                      ctypes.c_int32(len(to_list)), ctypes.c_int32(METRIC[metric]),
                                      ctypes.c_double(score_cutoff), ctypes.c_int32(int(exclude_self)),
                                      ctypes.c_int64(self_shift), _p(bi), _p(bs), _p(bd),
                                      ctypes.c_int32(n_threads))
    return bi, bs, bd
