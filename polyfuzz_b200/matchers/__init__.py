from ._base import BaseMatcher
from ._tfidf import TFIDF
from ._rapidfuzz import RapidFuzz, EditDistance
from ._embeddings import Embeddings

__all__ = ["BaseMatcher", "TFIDF", "RapidFuzz", "EditDistance", "Embeddings"]
