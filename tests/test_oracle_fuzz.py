"""The rapidfuzz-scorer oracle (oracle/fuzz.py) against the known answers rapidfuzz publishes
(tests/golden/rapidfuzz_published.json, sources inside), and the C edit-distance oracle against both."""
import json
import os

import numpy as np
import pytest

from oracle import fuzz
from oracle import native as onative


@pytest.fixture(scope="module")
def published(golden_dir):
    return json.load(open(os.path.join(golden_dir, "rapidfuzz_published.json")))


def test_published_scorer_vectors_bit_exact(published):
    assert len(published["scorers"]) >= 19
    for v in published["scorers"]:
        got = fuzz.SCORERS[v["fn"]](v["a"], v["b"])
        assert got == v["expect"], (v, got)                       # == on floats: all 15-17 published digits


def test_published_extract_one(published):
    for v in published["extract_one"]:
        r = fuzz.extract_one(v["query"], v["choices"], fuzz.SCORERS[v["scorer"]])
        assert list(r) == v["expect"]


def test_published_distances_pin_the_c_oracle(published):
    for v in published["distances"]:
        if v["fn"] == "levenshtein":
            assert int(onative.editdist_matrix([v["a"]], [v["b"]], "lev")[0, 0]) == v["expect"]
        elif v["fn"] == "indel":
            assert int(onative.editdist_matrix([v["a"]], [v["b"]], "indel")[0, 0]) == v["expect"]
            assert fuzz.indel_distance(v["a"], v["b"]) == v["expect"]
        else:
            _, s, _ = onative.editdist_argbest([v["a"]], [v["b"]], "norm_lev")
            assert s[0] == v["expect"]
    # fuzz.ratio of the C oracle == the restated rapidfuzz expression, bit for bit, on the published pairs and on random ones
    pairs = [(v["a"], v["b"]) for v in published["scorers"]]
    rng = np.random.default_rng(0)
    alpha = "abcde fgh"
    for _ in range(300):
        pairs.append(("".join(rng.choice(list(alpha), rng.integers(0, 14))), "".join(rng.choice(list(alpha), rng.integers(0, 14)))))
    for a, b in pairs:
        _, s, _ = onative.editdist_argbest([a], [b], "ratio", score_cutoff=float("-inf"))
        assert s[0] == fuzz.ratio(a, b), (a, b)


def test_scorer_properties():
    rng = np.random.default_rng(1)
    words = ["the", "of", "night", "day", "love", "man", "a", "x1", "last", "story", "Dead"]
    for _ in range(400):
        a = " ".join(rng.choice(words, rng.integers(1, 5))); b = " ".join(rng.choice(words, rng.integers(1, 7)))
        for fn in fuzz.SCORERS.values():
            s = fn(a, b)
            assert 0 <= s <= 100
            assert fn(a, a) == 100
        assert fuzz.ratio(a, b) == fuzz.ratio(b, a)
        assert fuzz.token_sort_ratio(a, b) == fuzz.token_sort_ratio(" ".join(reversed(a.split())), b)
        assert fuzz.partial_ratio(a, b) == fuzz.partial_ratio(b, a)
        assert fuzz.WRatio(a, b) >= fuzz.ratio(a, b)
        if set(a.split()) & set(b.split()):
            assert fuzz.partial_token_ratio(a, b) == 100
    assert fuzz.WRatio("", "abc") == 0 and fuzz.QRatio("", "") == 0 and fuzz.ratio("", "") == 100 and fuzz.partial_ratio("", "") == 100
    assert fuzz.extract_one("zzz", [], fuzz.WRatio) is None
    # score_cutoff semantics of extractOne: first maximal choice; a cutoff above every score gives None
    assert fuzz.extract_one("apple", ["apples", "apple", "apple"], fuzz.ratio) == ("apple", 100.0, 1)
    assert fuzz.extract_one("apple", ["mouse"], fuzz.ratio, score_cutoff=95) is None
