"""The plugin type.  When the reference package is importable its own `polyfuzz.models.BaseMatcher`
is used, so the B200 matchers ARE reference plugins (isinstance checks in polyfuzz/polyfuzz.py:127-151
pass and `PolyFuzz(method=TFIDF(...))` works unmodified).  Otherwise an identical ABC is defined
(mirror of polyfuzz/models/_base.py:6-31: abstract match(), attributes model_id and type)."""
from abc import ABC, abstractmethod
from typing import List

import pandas as pd

try:                                                    # pragma: no cover - depends on the environment
    from polyfuzz.models import BaseMatcher as _RefBaseMatcher
    BaseMatcher = _RefBaseMatcher
    REFERENCE_BASE = True
except Exception:                                       # reference not installed: same contract, own ABC
    REFERENCE_BASE = False

    class BaseMatcher(ABC):
        """The abstract BaseMatching to be modelled after for string matching"""

        def __init__(self, model_id: str = "Model 0"):
            self.model_id = model_id
            self.type = "Base Model"

        @abstractmethod
        def match(self, from_list: List[str], to_list: List[str] = None, **kwargs) -> pd.DataFrame:
            """Returns a DataFrame with columns From, To, Similarity (one row per from_list element)."""
            raise NotImplementedError()
