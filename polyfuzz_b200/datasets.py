"""Benchmark inputs of the reference (polyfuzz/datasets/_load_data.py:6-40 downloads them over HTTP; there is
no network here).  The two JSON files of the reference's `data/` directory travel with the repository as test
fixtures (tests/golden/data/); when they are absent the seeded synthetic stand-ins of `synth.py` are used and
the caller is told so.

    load_company_names() -> (list[str] of 100 000 names, "real" | "synthetic")
    load_movie_titles()  -> ({"Netflix": [...6 172], "IMDB": [...80 852]}, "real" | "synthetic")
"""
import json
import os

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "data")


def _read(name):
    path = os.path.join(_DATA, name)
    if not os.path.exists(path):
        return None
    with open(path, "r", encoding="utf-8") as f:
        return json.load(f)


def load_company_names(n=100_000, seed=0):
    data = _read("company_names.json")
    if data is not None and len(data) >= n:
        return list(data[:n]), "real"
    from . import synth
    return synth.company_names(n, seed=seed), "synthetic"


def load_movie_titles():
    data = _read("movie_titles.json")
    if data is not None:
        return {"Netflix": list(data["Netflix"]), "IMDB": list(data["IMDB"])}, "real"
    from . import synth
    return {"Netflix": synth.titles(6172, seed=1), "IMDB": synth.titles(80852, seed=2)}, "synthetic"
