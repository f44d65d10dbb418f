"""K3 host driver: all-pairs edit distance with per-row arg-best on the GPU (include/pfz.h, pfz_lev_*).

Semantics restated from the reference's call sites (polyfuzz/models/_rapidfuzz.py:99-113,
polyfuzz/models/_distance.py:89-102) and rapidfuzz's published definitions:
    "ratio"     fuzz.ratio           = (1 - indel/(|a|+|b|)) * 100            in [0, 100]
    "norm_lev"  Levenshtein.normalized_similarity = 1 - lev/max(|a|,|b|)       in [0, 1]
    "lev", "indel"  raw distances (best = smallest)
Best match of a from-string = first to-string (lowest index) with the maximal score >= score_cutoff.
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


def edit_argbest(from_list, to_list, metric="ratio", score_cutoff=0.0, exclude_self=False, self_shift=0,
                 want_matrix=False, n_splits=None):
    """Returns (best_idx int32[n_from] (-1 = none), best_score float64[n_from], best_dist int32[n_from]
    [, matrix int32[n_from, n_to]]) as device tensors."""
    dev = _dev()
    n_from, n_to = len(from_list), len(to_list)
    fblob, foff, _ = pack_strings(from_list)              # uint8 (ASCII list) or uint32 code points
    tblob, toff, _ = pack_strings(to_list)
    flens = np.diff(foff)
    if n_from and flens.max() > MAX_LEN:
        raise ValueError(f"from-string {int(flens.argmax())} has {int(flens.max())} code points; the edit-distance "
                         f"kernel supports at most {MAX_LEN}")
    best_idx = torch.full((max(n_from, 1),), -1, dtype=torch.int32, device=dev)
    best_score = torch.zeros(max(n_from, 1), dtype=torch.float64, device=dev)
    best_dist = torch.full((max(n_from, 1),), -1, dtype=torch.int32, device=dev)
    matrix = torch.zeros((max(n_from, 1), max(n_to, 1)), dtype=torch.int32, device=dev) if want_matrix else None
    if n_from == 0 or n_to == 0:
        out = (best_idx[:n_from], best_score[:n_from], best_dist[:n_from])
        return out + (matrix[:n_from, :n_to],) if want_matrix else out

    def blob_to_dev(b):
        if b.size == 0:
            return torch.zeros(1, dtype=torch.int32, device=dev)
        return _to_dev(b).to(torch.int32) if b.dtype == np.uint8 else _to_dev(b.view(np.int32), torch.int32)
    d_fblob = blob_to_dev(fblob)
    d_foff = _to_dev(foff)
    d_tblob = blob_to_dev(tblob)
    d_toff = _to_dev(toff)
    # to-strings sorted by length, groups of 32, 4 symbols per word
    tlens = np.diff(toff)
    order = np.argsort(tlens, kind="stable").astype(np.int32)
    n_grp = (n_to + 31) // 32
    gmax = tlens[order[np.minimum(np.arange(n_grp) * 32 + 31, n_to - 1)]]
    gwords = ((gmax + 3) // 4) * 32
    goff = np.zeros(n_grp + 1, dtype=np.int64); np.cumsum(gwords, out=goff[1:])
    d_order = _to_dev(order); d_goff = _to_dev(goff)
    packed = torch.empty(max(int(goff[-1]), 1), dtype=torch.int32, device=dev)
    slen = torch.empty(n_to, dtype=torch.int32, device=dev)

    if n_splits is None:
        # ~4 (pattern, to-split) tasks per resident warp: patterns differ in length, finer tasks balance the tail
        want = 4 * 148 * 48
        n_splits = max(1, min(n_grp, (want + n_from - 1) // n_from))
    n_splits = max(1, min(int(n_splits), n_grp))
    part_idx = torch.full((n_splits, n_from), -1, dtype=torch.int32, device=dev)
    part_score = torch.zeros((n_splits, n_from), dtype=torch.float64, device=dev)
    part_dist = torch.full((n_splits, n_from), -1, dtype=torch.int32, device=dev)
    counter = torch.zeros(n_splits, dtype=torch.int32, device=dev)
    classes = np.array([_word_class(int(m)) for m in flens], dtype=np.int32) if n_from < 4096 else \
        np.select([flens <= 32, flens <= 64, flens <= 128, flens <= 256, flens <= 512], [0, 1, 2, 4, 8], 16).astype(np.int32)

    keep = []
    for lo, hi in _alphabet_batches(fblob, foff):
        cps = np.unique(fblob[foff[lo]:foff[hi]])
        table = np.zeros(N_CODE_POINTS, dtype=np.uint8)
        cps = cps.astype(np.int64)
        table[cps[cps < N_CODE_POINTS]] = np.arange(1, len(cps) + 1, dtype=np.uint8)[:int((cps < N_CODE_POINTS).sum())]
        d_table = _to_dev(table)
        _lib.call("pfz_lev_pack", _p(d_tblob), _p(d_toff), _p(d_order), n_to, _p(d_table), _p(d_goff), _p(packed), _p(slen), _stream())
        for nw in (0, 1, 2, 4, 8, 16):
            ids = np.nonzero(classes[lo:hi] == nw)[0].astype(np.int32) + lo
            if len(ids) == 0:
                continue
            d_ids = _to_dev(ids); keep.append(d_ids)
            _lib.call("pfz_lev_argbest", _p(d_fblob), _p(d_foff), n_from, _p(d_ids), len(ids), nw, _p(d_table), _p(packed),
                      _p(d_goff), _p(slen), _p(d_order), n_to, METRIC[metric], float(score_cutoff), int(bool(exclude_self)),
                      int(self_shift), n_splits, _p(part_idx), _p(part_score), _p(part_dist), _p(matrix),
                      int(matrix.stride(0)) if matrix is not None else 0, _p(counter), _stream())
        keep.append(d_table)
    _lib.call("pfz_lev_merge", _p(part_idx), _p(part_score), _p(part_dist), n_splits, n_from, _p(best_idx), _p(best_score),
              _p(best_dist), _stream())
    out = (best_idx[:n_from], best_score[:n_from], best_dist[:n_from])
    return out + (matrix[:n_from, :n_to],) if want_matrix else out
