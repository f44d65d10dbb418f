"""RapidFuzz / EditDistance matchers -- drop-ins for polyfuzz.models.RapidFuzz
(polyfuzz/models/_rapidfuzz.py:11-113) and polyfuzz.models.EditDistance
(polyfuzz/models/_distance.py:12-102) whose all-pairs scoring runs on the GPU (K3 / K3b).

Scorers (rapidfuzz 3.x definitions, no preprocessing -- oracle/fuzz.py restates them and is pinned on rapidfuzz's
published known answers):
  * "WRatio" (the reference's default for RapidFuzz, fuzz.WRatio, _rapidfuzz.py:48), "QRatio", "partial_ratio",
    "token_sort_ratio", "token_set_ratio", "token_ratio", "partial_token_sort_ratio", "partial_token_set_ratio",
    "partial_token_ratio"                                         -> csrc/pfz_fuzz.cu (K3b)
  * "ratio" (fuzz.ratio, the reference's default for EditDistance, _distance.py:32) and "levenshtein"
    (Levenshtein.normalized_similarity)                                  -> csrc/pfz_lev.cu (K3)
A scorer may be given by name or as the rapidfuzz callable of that __name__.  Arbitrary Python callables cannot be
compiled to the device and raise NotImplementedError -- there is no CPU fallback.
Deviations from the reference, both documented reference bugs: a self-match excludes
index i only (the reference mutates the shared to_list, _rapidfuzz.py:103-104), and the matcher can be
reused for a two-list call after a self-match (`equal_lists` is per call)."""
from typing import Callable, List, Union

import numpy as np
import pandas as pd

from ._base import BaseMatcher
from .. import editdist, fuzzy
from ..distributed import get_comm, shard_bounds

_NAMES = {"ratio": "ratio", "levenshtein": "norm_lev", "norm_lev": "norm_lev", "normalized_similarity": "norm_lev",
          "normalized_levenshtein": "norm_lev"}
_FUZZ = {k.lower(): k for k in fuzzy.SCORER if k != "ratio"}


def _resolve_scorer(scorer, default) -> str:
    """-> "ratio" | "norm_lev" (K3) or one of fuzzy.SCORER (K3b)."""
    if scorer is None:
        scorer = default
    key = scorer.lower() if isinstance(scorer, str) else getattr(scorer, "__name__", "").lower()
    if key in _NAMES:
        return _NAMES[key]
    if key in _FUZZ:
        return _FUZZ[key]
    raise NotImplementedError(f"scorer {scorer!r} has no GPU implementation (supported: 'ratio', 'levenshtein', "
                              f"{sorted(_FUZZ.values())}); polyfuzz_b200 has no CPU fallback")


def _argbest(from_list, targets, metric, cutoff, self_match, distributed):
    """Best to-string per from-string on this GPU, or -- distributed=True under torchrun -- on the to_list row-block of every
    rank followed by ONE all-gather of the per-shard bests and the canonical merge (score desc, global index asc): all ranks
    get the single-GPU result (the reference's own fan-out is per from-row, polyfuzz/models/_rapidfuzz.py:92-95)."""
    comm = get_comm() if distributed else None
    token_scorer = metric in fuzzy.SCORER and metric != "ratio"
    if comm is None:
        if token_scorer:
            return fuzzy.fuzz_argbest(from_list, targets, metric, cutoff, exclude_self=self_match) + (None,)
        return editdist.edit_argbest(from_list, targets, metric, cutoff, exclude_self=self_match)
    lo, hi = shard_bounds(len(targets), comm.world_size, comm.rank)
    if token_scorer:
        bi, bs = fuzzy.fuzz_argbest(from_list, targets[lo:hi], metric, cutoff, exclude_self=self_match, self_shift=-lo, to_index_base=lo)
        bd = torch_full_like_int(bi)
    else:
        Q = editdist.EditQueries(from_list)
        T = editdist.EditTargets(targets[lo:hi])
        bi, bs, bd = editdist.edit_argbest_staged(Q, T, metric, cutoff, exclude_self=self_match, self_shift=-lo, to_index_base=lo)
    gi, gs, gd = comm.all_gather_best(bi, bs, bd)
    return editdist.lev_merge(gi, gs, gd)


def torch_full_like_int(t):
    import torch
    return torch.full_like(t, -1)

class RapidFuzz(BaseMatcher):
    """Edit-distance matcher (GPU).  Arguments as in the reference: n_jobs (accepted, ignored -- the GPU
    scores all pairs in one launch), score_cutoff in [0,1], scorer (default fuzz.WRatio, as the reference), model_id."""

    def __init__(self, n_jobs: int = 1, score_cutoff: float = 0, scorer: Union[str, Callable] = "WRatio", model_id: str = None,
                 distributed: bool = False):
        super().__init__(model_id)
        self.type = "EditDistance"
        self.distributed = distributed
        self.score_cutoff = score_cutoff * 100
        self.scorer = scorer
        self._metric = _resolve_scorer(scorer, "WRatio")
        self.equal_lists = False
        self.n_jobs = n_jobs

    def match(self, from_list: List[str], to_list: List[str] = None, **kwargs) -> pd.DataFrame:
        """(from, best to, score/100); no candidate with score >= score_cutoff -> (from, None, 0.0)
        (polyfuzz/models/_rapidfuzz.py:106-113)."""
        self_match = to_list is None
        targets = from_list if self_match else to_list
        scale = 1.0 if self._metric == "norm_lev" else 100.0
        cutoff = self.score_cutoff / 100.0 if self._metric == "norm_lev" else self.score_cutoff
        idx, score, _ = _argbest(from_list, targets, self._metric, cutoff, self_match, self.distributed)
        idx = idx.cpu().numpy(); score = score.cpu().numpy() / scale
        to_arr = np.empty(len(targets) + 1, dtype=object); to_arr[:-1] = targets; to_arr[-1] = None
        sel = np.where(idx >= 0, idx, len(targets))
        return pd.DataFrame({"From": pd.Series(list(from_list), dtype=object), "To": pd.Series(to_arr[sel], dtype=object),
                             "Similarity": np.where(idx >= 0, score, 0.0)})


class EditDistance(BaseMatcher):
    """Edit-distance matcher with the reference's EditDistance surface (n_jobs, scorer, model_id, normalize):
    Similarity is the scorer's raw value (fuzz.ratio: 0..100) of the best to-string, min-max normalised
    over the column when `normalize` (polyfuzz/models/_distance.py:83-86)."""

    def __init__(self, n_jobs: int = 1, scorer: Union[str, Callable] = "ratio", model_id: str = None, normalize: bool = True,
                 distributed: bool = False):
        super().__init__(model_id)
        self.type = "EditDistance"
        self.distributed = distributed
        self.scorer = scorer
        self._metric = _resolve_scorer(scorer, "ratio")
        self.normalize = normalize
        self.equal_lists = False
        self.n_jobs = n_jobs

    def match(self, from_list: List[str], to_list: List[str] = None, **kwargs) -> pd.DataFrame:
        self_match = to_list is None
        targets = from_list if self_match else to_list
        if len(targets) - (1 if self_match else 0) < 1:
            raise ValueError("attempt to get argmax of an empty sequence")         # np.argmax on [] in the reference
        # np.argmax over the scorer's values (polyfuzz/models/_distance.py:98-99): no cutoff; the token scorers take
        # score_cutoff = 0 (their default), which every score passes
        no_cut = 0.0 if self._metric in fuzzy.SCORER and self._metric != "ratio" else float("-inf")
        idx, score, _ = _argbest(from_list, targets, self._metric, no_cut, self_match, self.distributed)
        idx = idx.cpu().numpy(); score = score.cpu().numpy()
        to_arr = np.empty(len(targets), dtype=object); to_arr[:] = targets
        matches = pd.DataFrame({"From": pd.Series(list(from_list), dtype=object), "To": pd.Series(to_arr[idx], dtype=object),
                                "Similarity": score})
        if self.normalize:
            s = matches["Similarity"]
            matches["Similarity"] = (s - s.min()) / (s.max() - s.min())
        return matches
