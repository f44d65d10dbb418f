"""polyfuzz_b200 -- B200-native (sm_100a) pairwise string-similarity hot path behind PolyFuzz's
BaseMatcher plugin API.  See DESIGN.md / INTEGRATION.md."""
from .matchers import BaseMatcher, TFIDF, RapidFuzz, EditDistance, Embeddings  # noqa: F401

__version__ = "0.1.0"
