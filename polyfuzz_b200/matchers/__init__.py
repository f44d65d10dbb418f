from ._base import BaseMatcher
from ._tfidf import TFIDF

__all__ = ["BaseMatcher", "TFIDF"]
