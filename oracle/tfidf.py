"""oracle/tfidf.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's char-n-gram TF-IDF vectoriser:
    polyfuzz/models/_tfidf.py:142-146   _clean_string
    polyfuzz/models/_tfidf.py:120-139   TFIDF._create_ngrams
    polyfuzz/models/_tfidf.py:102-118   TFIDF._extract_tf_idf -> sklearn TfidfVectorizer(min_df=1, analyzer=...)
and of the scikit-learn arithmetic it delegates to (scikit-learn is the installed third-party
package that executes the reference's arithmetic; file:line into sklearn 1.9.0):
    feature_extraction/text.py:1257-1320  _count_vocab   (tf = raw count, alphabetical vocabulary)
    feature_extraction/text.py:1651-1696  idf = ln((1+N)/(1+df)) + 1
    feature_extraction/text.py:1698-1739  x = tf*idf ; row-wise l2 normalise
    utils/sparsefuncs_fast.pyx:578-605    norm = sqrt(sum x*x in column order) ; x /= norm

Two independent statements are provided and cross-checked in tests/test_oracle_*.py:
  * `fit_transform_sklearn`  -- the restated analyzer fed to the real TfidfVectorizer;
  * `TfidfOracle`            -- a from-scratch numpy statement (no sklearn).
Parity is PINNED: both are checked against the unmodified reference through tests/golden/.
"""
import re
import numpy as np
import scipy.sparse as sp

_NON_ALNUM = re.compile(r"[^A-Za-z0-9 ]+")
_SPACES = re.compile(r"\s+")


def clean_string(s: str) -> str:
    """_tfidf.py:142-146: lower-case, delete everything outside [A-Za-z0-9 ], collapse whitespace
    runs to one space, strip."""
    s = _NON_ALNUM.sub("", s.lower())
    return _SPACES.sub(" ", s).strip()


def create_ngrams(s: str, n_gram_range=(3, 3), clean=True, remove_space_ngrams=True):
    """_tfidf.py:120-139: for n in lo..hi, all length-n sliding windows; optionally drop windows
    containing a space."""
    if clean:
        s = clean_string(s)
    out = []
    for n in range(n_gram_range[0], n_gram_range[1] + 1):
        for i in range(len(s) - n + 1):
            g = s[i:i + n]
            if remove_space_ngrams and " " in g:
                continue
            out.append(g)
    return out


def fit_transform_sklearn(from_list, to_list=None, n_gram_range=(3, 3), clean=True, remove_space_ngrams=True):
    """_tfidf.py:102-118 with the restated analyzer.  Returns (tf_idf_from, tf_idf_to, vectorizer)."""
    from sklearn.feature_extraction.text import TfidfVectorizer
    an = lambda s: create_ngrams(s, n_gram_range, clean, remove_space_ngrams)  # noqa: E731
    if to_list:
        vec = TfidfVectorizer(min_df=1, analyzer=an).fit(list(to_list) + list(from_list))
        tf_to = vec.transform(to_list)
        tf_from = vec.transform(from_list)
    else:
        vec = TfidfVectorizer(min_df=1, analyzer=an).fit(from_list)
        tf_to = vec.transform(from_list)
        tf_from = tf_to
    return tf_from, tf_to, vec


class TfidfOracle:
    """From-scratch numpy statement of the fitted vectoriser (vocabulary_, idf_, transform)."""

    def __init__(self, n_gram_range=(3, 3), clean=True, remove_space_ngrams=True):
        self.n_gram_range = tuple(n_gram_range)
        self.clean = clean
        self.remove_space_ngrams = remove_space_ngrams
        self.vocabulary = None      # sorted list of n-gram strings (column j <-> vocabulary[j])
        self.idf = None

    def _analyze(self, s):
        return create_ngrams(s, self.n_gram_range, self.clean, self.remove_space_ngrams)

    def fit(self, corpus):
        df = {}
        n = 0
        for doc in corpus:
            n += 1
            for g in set(self._analyze(doc)):
                df[g] = df.get(g, 0) + 1
        if not df:
            raise ValueError("empty vocabulary; perhaps the documents only contain stop words")
        self.vocabulary = sorted(df)                      # Python str order == code point order
        self._col = {g: j for j, g in enumerate(self.vocabulary)}
        dfv = np.array([df[g] for g in self.vocabulary], dtype=np.float64)
        # sklearn: idf = np.log((n_samples + 1) / (df + 1)) + 1   (smooth_idf=True)
        self.idf = np.log((n + 1.0) / (dfv + 1.0)) + 1.0
        self.df = dfv.astype(np.int64)
        return self

    def transform(self, docs):
        indptr = [0]
        indices = []
        data = []
        for doc in docs:
            cnt = {}
            for g in self._analyze(doc):
                j = self._col.get(g)
                if j is not None:
                    cnt[j] = cnt.get(j, 0) + 1
            cols = sorted(cnt)
            x = np.array([cnt[j] for j in cols], dtype=np.float64) * self.idf[cols] if cols else np.zeros(0)
            ss = 0.0
            for v in x:                                     # column order, product rounded then added
                ss = ss + v * v
            if ss > 0.0:
                x = x / np.sqrt(ss)
            indices.extend(cols)
            data.extend(x.tolist())
            indptr.append(len(indices))
        return sp.csr_matrix((np.array(data, dtype=np.float64), np.array(indices, dtype=np.int32),
                              np.array(indptr, dtype=np.int32)), shape=(len(indptr) - 1, len(self.vocabulary)))
