"""TFIDF matcher -- drop-in for polyfuzz.models.TFIDF (polyfuzz/models/_tfidf.py:11-146) running the
vectoriser (K1) and the sparse cosine top-n (K2) on a B200."""
from typing import List, Tuple

import numpy as np
import pandas as pd

from ._base import BaseMatcher
from ._utils import assemble_matches, clip_top_n
from .. import engine


class TFIDF(BaseMatcher):
    """
    A character based n-gram TF-IDF to approximate edit distance (same constructor as the reference).

    Arguments:
        n_gram_range: The n_gram_range on a character-level
        clean_string: Whether to clean the string such that only alphanumerical characters are kept
        min_similarity: The minimum similarity between strings, otherwise return 0 similarity
        top_n: The number of matches you want returned
        cosine_method: accepted for API compatibility ("sparse" | "sklearn" | "knn"); every value runs
                       the same fused sparse kernel, which implements the `sparse` branch semantics
                       (candidates need score > min_similarity, polyfuzz/models/_utils.py:82).
        model_id: The name of the particular instance, used when comparing models
        remove_space_ngrams: Remove n-grams that contain a space
    """

    def __init__(self,
                 n_gram_range: Tuple[int, int] = (3, 3),
                 clean_string: bool = True,
                 min_similarity: float = 0.75,
                 top_n: int = 1,
                 cosine_method: str = "sparse",
                 model_id: str = None,
                 remove_space_ngrams=True):
        super().__init__(model_id)
        self.type = "TF-IDF"
        self.n_gram_range = n_gram_range
        self.clean_string = clean_string
        self.min_similarity = min_similarity
        self.cosine_method = cosine_method
        self.top_n = top_n
        self.vectorizer = None
        self.tf_idf_to = None
        self.remove_space_ngrams = remove_space_ngrams
        self._index = None              # device inverted index of tf_idf_to (rebuilt lazily after unpickling)

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_index"] = None
        if self.tf_idf_to is not None and not hasattr(self.tf_idf_to, "tocsr"):
            st["tf_idf_to"] = self.tf_idf_to.to_scipy()          # device CSR -> host scipy for joblib.dump
        return st

    def match(self, from_list: List[str], to_list: List[str] = None, re_train: bool = True) -> pd.DataFrame:
        """Match two lists of strings to each other and return the most similar strings
        (polyfuzz/models/_tfidf.py:68-100)."""
        top_idx, top_val, top_n = self.match_arrays(from_list, to_list, re_train)
        return assemble_matches(from_list, to_list, top_idx.cpu().numpy(), top_val.cpu().numpy())

    def match_arrays(self, from_list, to_list=None, re_train=True):
        """Device-side result: (top_idx int32[n,k] with -1 for no match, top_val float64[n,k], k)."""
        tf_idf_from, tf_idf_to = self._extract_tf_idf(from_list, to_list, re_train)
        top_n = clip_top_n(self.top_n, to_list)
        if top_n < 1:
            raise ValueError("top_n must be >= 1 and to_list must not be empty")
        idx, val = engine.spcos_topk(tf_idf_from, self._index, top_n, self.min_similarity,
                                     self_match=to_list is None)
        return idx, val, top_n

    def _extract_tf_idf(self, from_list, to_list=None, re_train=True):
        """polyfuzz/models/_tfidf.py:102-118 (note `if to_list:` -- an empty list means self-match for the
        vectoriser while `cosine_similarity` still tests `is not None`)."""
        if re_train:
            self.vectorizer = engine.NgramTfidf(self.n_gram_range, self.clean_string, self.remove_space_ngrams)
        elif self.vectorizer is None:
            raise ValueError("re_train=False needs a fitted model (call match/fit first)")
        vec = self.vectorizer
        if to_list:
            if re_train:
                rows_to, rows_from = vec.fit_rows([to_list, from_list])
                self.tf_idf_to = vec.emit(rows_to)
                self._index = None
            else:
                rows_from = vec.rows(from_list)
            tf_idf_from = vec.emit(rows_from)
        else:
            if re_train:
                (rows_from,) = vec.fit_rows([from_list])
                self.tf_idf_to = vec.emit(rows_from)
                self._index = None
            tf_idf_from = self._device_to()
        if self._index is None:
            self._index = engine.SparseIndex(self._device_to())
        return tf_idf_from, self.tf_idf_to

    def _device_to(self):
        if hasattr(self.tf_idf_to, "tocsr"):                      # restored from a pickle
            self.tf_idf_to = engine.CsrMatrix.from_scipy(self.tf_idf_to)
        return self.tf_idf_to
