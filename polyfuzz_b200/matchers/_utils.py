"""Result assembly shared by the matchers -- the tail of polyfuzz/models/_utils.py:104-125, built
column-wise from the top-k arrays instead of through a (1+2k) x n unicode ndarray."""
from typing import List, Optional

import numpy as np
import pandas as pd


def clip_top_n(top_n: int, to_list: Optional[List[str]]) -> int:
    """polyfuzz/models/_utils.py:54-56 -- only when a to_list is given."""
    if to_list is not None:
        top_n = min(top_n, len(set(to_list)))
    return top_n


def assemble_matches(from_list, to_list, top_idx: np.ndarray, top_val: np.ndarray) -> pd.DataFrame:
    """top_idx int32[n,k] (global to-index, -1 = none), top_val float64[n,k] (unrounded scores).
    Columns From, To, Similarity, To_2, Similarity_2, ... ; similarities rounded to 3 decimals
    (_utils.py:102); Similarity < 0.001 -> 0.0 and To -> None (_utils.py:119-123)."""
    if to_list is None:
        to_list = from_list
    n, k = top_idx.shape
    to_arr = np.empty(len(to_list) + 1, dtype=object)
    to_arr[:-1] = to_list
    to_arr[-1] = None
    cols = {"From": list(from_list)}
    for r in range(k):
        sims = np.round(top_val[:, r], 3)
        idx = top_idx[:, r].astype(np.int64)
        low = (sims < 0.001) | (idx < 0)
        sims = np.where(low, 0.0, sims)
        idx = np.where(low, len(to_list), idx)
        cols["To" if r == 0 else f"To_{r + 1}"] = to_arr[idx]
        cols["Similarity" if r == 0 else f"Similarity_{r + 1}"] = sims
    return pd.DataFrame(cols)
