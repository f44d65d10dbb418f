"""K3 host driver: all-pairs edit distance with per-row arg-best on the GPU (include/pfz.h, pfz_lev_*).

Semantics restated from the reference's call sites (polyfuzz/models/_rapidfuzz.py:99-113,
polyfuzz/models/_distance.py:89-102) and rapidfuzz's published definitions:
    "ratio"     fuzz.ratio           = (1 - indel/(|a|+|b|)) * 100            in [0, 100]
    "norm_lev"  Levenshtein.normalized_similarity = 1 - lev/max(|a|,|b|)       in [0, 1]
    "lev", "indel"  raw distances (best = smallest)
Best match of a from-string = first to-string (lowest index) with the maximal score >= score_cutoff.

Two levels: `EditQueries` / `EditTargets` stage a from-list / to-list in HBM once (host packing, length sort,
alphabet batches); `edit_argbest_staged` only enqueues kernels, so a staged pair can be scored repeatedly
(bench.py's device-timed leg, the multi-GPU shards) without touching the host lists again.
"""
import numpy as np
import torch

from . import _lib
from .engine import _dev, _p, _stream, _to_dev
from .strings import pack_strings

METRIC = {"lev": 0, "indel": 1, "norm_lev": 2, "ratio": 3}
N_CODE_POINTS = 0x110000
MAX_LEN = 1024


def _word_class(m):
    """n_words argument of pfz_lev_argbest for a pattern of m code points."""
    if m <= 32:
        return 0
    for nw in (1, 2, 4, 8, 16):
        if m <= 64 * nw:
            return nw
    raise ValueError(f"from-string of {m} code points exceeds the supported maximum of {MAX_LEN}")


def _alphabet_batches(blob, offsets):
    """Split the from-rows into consecutive batches whose joint alphabet has <= 255 code points."""
    n = len(offsets) - 1
    if n == 0:
        return []
    if len(np.unique(blob)) <= 255:
        return [(0, n)]
    batches, lo, cur = [], 0, set()
    for i in range(n):
        cps = set(np.unique(blob[offsets[i]:offsets[i + 1]]).tolist())
        if len(cps) > 255:
            raise ValueError(f"from-string {i} has more than 255 distinct code points")
        if len(cur | cps) > 255:
            batches.append((lo, i)); lo, cur = i, set()
        cur |= cps
    batches.append((lo, n))
    return batches


def _blob_to_dev(b):
    if b.size == 0:
        return torch.zeros(1, dtype=torch.int32, device=_dev())
    return _to_dev(b).to(torch.int32) if b.dtype == np.uint8 else _to_dev(b.view(np.int32), torch.int32)


class EditQueries:
    """A from-list staged in HBM: UTF-32 blob + offsets, the rows of every word class, and per alphabet batch
    the code-point -> byte-symbol table (the kernel's symbols are bytes, see pfz.h)."""

    def __init__(self, from_list):
        self.n = len(from_list)
        blob, off, _ = pack_strings(from_list)
        lens = np.diff(off)
        if self.n and lens.max() > MAX_LEN:
            raise ValueError(f"from-string {int(lens.argmax())} has {int(lens.max())} code points; the edit-distance "
                             f"kernel supports at most {MAX_LEN}")
        self.lens = lens
        self.n_chars = int(blob.size)
        self.h2d_bytes = blob.nbytes + off.nbytes
        if self.n == 0:
            self.batches = []
            return
        self.d_blob = _blob_to_dev(blob)
        self.d_off = _to_dev(off)
        classes = np.select([lens <= 32, lens <= 64, lens <= 128, lens <= 256, lens <= 512], [0, 1, 2, 4, 8], 16).astype(np.int32)
        self.batches = []                                   # [(d_table, [(n_words, d_ids, n_ids), ...])]
        for lo, hi in _alphabet_batches(blob, off):
            cps = np.unique(blob[off[lo]:off[hi]]).astype(np.int64)
            table = np.zeros(N_CODE_POINTS, dtype=np.uint8)
            ok = cps < N_CODE_POINTS
            table[cps[ok]] = np.arange(1, len(cps) + 1, dtype=np.uint8)[:int(ok.sum())]
            groups = []
            for nw in (0, 1, 2, 4, 8, 16):
                ids = np.nonzero(classes[lo:hi] == nw)[0].astype(np.int32) + lo
                if len(ids):
                    groups.append((nw, _to_dev(ids), len(ids)))
                    self.h2d_bytes += ids.nbytes
            self.batches.append((_to_dev(table), groups))
            self.h2d_bytes += table.nbytes


class EditTargets:
    """A to-list (or one row-block shard of it) staged in HBM: sorted by length, groups of 32, transposed,
    4 byte-symbols per 32-bit word (filled per alphabet batch by pfz_lev_pack)."""

    def __init__(self, to_list):
        self.n = len(to_list)
        blob, off, _ = pack_strings(to_list)
        self.n_chars = int(blob.size)
        self.h2d_bytes = blob.nbytes + off.nbytes
        if self.n == 0:
            return
        dev = _dev()
        self.d_blob = _blob_to_dev(blob)
        self.d_off = _to_dev(off)
        tlens = np.diff(off)
        self.lens = tlens
        order = np.argsort(tlens, kind="stable").astype(np.int32)
        self.n_grp = (self.n + 31) // 32
        gmax = tlens[order[np.minimum(np.arange(self.n_grp) * 32 + 31, self.n - 1)]]
        gwords = ((gmax + 3) // 4) * 32
        goff = np.zeros(self.n_grp + 1, dtype=np.int64); np.cumsum(gwords, out=goff[1:])
        self.d_order = _to_dev(order); self.d_goff = _to_dev(goff)
        self.h2d_bytes += order.nbytes + goff.nbytes
        self.packed = torch.empty(max(int(goff[-1]), 1), dtype=torch.int32, device=dev)
        self.slen = torch.empty(self.n, dtype=torch.int32, device=dev)


def default_splits(n_from, n_grp):
    # ~4 (pattern, to-split) tasks per resident warp: patterns differ in length, finer tasks balance the tail
    want = 4 * 148 * 48
    return max(1, min(n_grp, (want + max(n_from, 1) - 1) // max(n_from, 1)))


def edit_argbest_staged(Q, T, metric="ratio", score_cutoff=0.0, exclude_self=False, self_shift=0, want_matrix=False,
                        n_splits=None, to_index_base=0):
    """Kernels only.  Returns (best_idx int32[n_from] (-1 = none; + to_index_base otherwise), best_score float64[n_from],
    best_dist int32[n_from] [, matrix int32[n_from, n_to]]) as device tensors."""
    dev = _dev()
    n_from, n_to = Q.n, T.n
    best_idx = torch.full((max(n_from, 1),), -1, dtype=torch.int32, device=dev)
    best_score = torch.zeros(max(n_from, 1), dtype=torch.float64, device=dev)
    best_dist = torch.full((max(n_from, 1),), -1, dtype=torch.int32, device=dev)
    matrix = torch.zeros((max(n_from, 1), max(n_to, 1)), dtype=torch.int32, device=dev) if want_matrix else None
    if n_from == 0 or n_to == 0:
        out = (best_idx[:n_from], best_score[:n_from], best_dist[:n_from])
        return out + (matrix[:n_from, :n_to],) if want_matrix else out
    if n_splits is None:
        n_splits = default_splits(n_from, T.n_grp)
    n_splits = max(1, min(int(n_splits), T.n_grp))
    part_idx = torch.full((n_splits, n_from), -1, dtype=torch.int32, device=dev)
    part_score = torch.zeros((n_splits, n_from), dtype=torch.float64, device=dev)
    part_dist = torch.full((n_splits, n_from), -1, dtype=torch.int32, device=dev)
    counter = torch.zeros(n_splits, dtype=torch.int32, device=dev)
    for d_table, groups in Q.batches:
        _lib.call("pfz_lev_pack", _p(T.d_blob), _p(T.d_off), _p(T.d_order), n_to, _p(d_table), _p(T.d_goff), _p(T.packed), _p(T.slen), _stream())
        for nw, d_ids, n_ids in groups:
            _lib.call("pfz_lev_argbest", _p(Q.d_blob), _p(Q.d_off), n_from, _p(d_ids), n_ids, nw, _p(d_table), _p(T.packed),
                      _p(T.d_goff), _p(T.slen), _p(T.d_order), n_to, METRIC[metric], float(score_cutoff), int(bool(exclude_self)),
                      int(self_shift), n_splits, _p(part_idx), _p(part_score), _p(part_dist), _p(matrix),
                      int(matrix.stride(0)) if matrix is not None else 0, _p(counter), _stream())
    _lib.call("pfz_lev_merge", _p(part_idx), _p(part_score), _p(part_dist), n_splits, n_from, _p(best_idx), _p(best_score),
              _p(best_dist), _stream())
    if to_index_base:
        best_idx = torch.where(best_idx >= 0, best_idx + int(to_index_base), best_idx)
    out = (best_idx[:n_from], best_score[:n_from], best_dist[:n_from])
    return out + (matrix[:n_from, :n_to],) if want_matrix else out


def edit_argbest(from_list, to_list, metric="ratio", score_cutoff=0.0, exclude_self=False, self_shift=0,
                 want_matrix=False, n_splits=None):
    """Host lists in, device tensors out (see edit_argbest_staged)."""
    _dev()
    Q = EditQueries(from_list)
    T = EditTargets(to_list)
    return edit_argbest_staged(Q, T, metric, score_cutoff, exclude_self, self_shift, want_matrix, n_splits)


def lev_merge(part_idx, part_score, part_dist):
    """[n_lists, n_from] partial bests (GLOBAL indices) -> the canonical best per row (score desc, index asc)."""
    n_lists, n_from = part_idx.shape
    dev = part_idx.device
    bi = torch.full((max(n_from, 1),), -1, dtype=torch.int32, device=dev)
    bs = torch.zeros(max(n_from, 1), dtype=torch.float64, device=dev)
    bd = torch.full((max(n_from, 1),), -1, dtype=torch.int32, device=dev)
    _lib.call("pfz_lev_merge", _p(part_idx.contiguous()), _p(part_score.contiguous()), _p(part_dist.contiguous()), n_lists, n_from,
              _p(bi), _p(bs), _p(bd), _stream())
    return bi[:n_from], bs[:n_from], bd[:n_from]
