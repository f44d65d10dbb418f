"""oracle/fuzz.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Pure-Python restatement of the rapidfuzz scorers the reference's edit-distance matchers call:
    polyfuzz/models/_rapidfuzz.py:48      scorer=fuzz.WRatio  (default of RapidFuzz and of PolyFuzz("EditDistance"),
                                          polyfuzz/polyfuzz.py:128-130)
    polyfuzz/models/_rapidfuzz.py:24-37   documented scorer list (ratio, partial_ratio, token_sort_ratio, token_set_ratio,
                                          token_ratio, partial_token_*_ratio, WRatio, QRatio)
    polyfuzz/models/_rapidfuzz.py:106-108 process.extractOne(query, choices, scorer=..., score_cutoff=...)
    polyfuzz/models/_distance.py:32,98    scorer=fuzz.ratio ; [scorer(a, b) for b in to_list] ; np.argmax

rapidfuzz is a third-party dependency that is absent from /root/reference and cannot be installed here (pin in the
reference: rapidfuzz>=0.13.1, setup.py:20; restated: the 3.x line, where fuzz.* and process.* apply NO preprocessing by
default, processor=None).  The algorithms below follow rapidfuzz's published pure-Python implementation
(src/rapidfuzz/fuzz_py.py, process_py.py, distance/Indel_py.py), including the order of the floating-point
operations -- `ratio` is (1 - dist/lensum) * 100 while the token-set parts use 100 - 100*dist/lensum -- and the way
score_cutoff is threaded through WRatio.

PARITY PIN: tests/golden/rapidfuzz_published.json holds the known answers rapidfuzz publishes in its README /
documentation (source URLs in the file); tests/test_oracle_fuzz.py checks every function here against them.  The
15-17 significant digits of those answers discriminate between the two float expressions above, the window set of
partial_ratio (prefix / full / suffix windows: "cowboys" vs "Dallas Cowboys" = 83.07692307692308) and WRatio's scale
factors.  One choice is NOT pinned by a published vector: WRatio uses the 0.9 partial scale for len_ratio <= 8.0
(fuzzywuzzy-compatible; `< 8.0` would differ only for length ratios of exactly 8).
"""
from math import ceil


# ---- Indel / LCS ----------------------------------------------------------------------------------------------
def lcs_len(a, b):
    """Longest common subsequence length (textbook DP)."""
    if not a or not b:
        return 0
    prev = [0] * (len(b) + 1)
    for ca in a:
        cur = [0]
        for j, cb in enumerate(b, 1):
            cur.append(prev[j - 1] + 1 if ca == cb else (prev[j] if prev[j] >= cur[j - 1] else cur[j - 1]))
        prev = cur
    return prev[-1]


def indel_distance(a, b):
    return len(a) + len(b) - 2 * lcs_len(a, b)


def _indel_norm_sim(a, b, score_cutoff=0.0):
    """Indel.normalized_similarity with its cutoff semantics: 1 - dist/(|a|+|b|), 0 when below the cutoff."""
    maximum = len(a) + len(b)
    norm_dist = indel_distance(a, b) / maximum if maximum else 0.0
    norm_sim = 1.0 - norm_dist
    return norm_sim if norm_sim >= score_cutoff else 0.0


def _norm_distance(dist, lensum, score_cutoff):
    score = (100 - 100 * dist / lensum) if lensum else 100
    return score if score >= score_cutoff else 0


def _split(s):
    return s.split()


def _join(tokens):
    return " ".join(tokens)


# ---- scorers (processor=None) -----------------------------------------------------------------------------------
def ratio(s1, s2, score_cutoff=0):
    return _indel_norm_sim(s1, s2, score_cutoff / 100) * 100


def QRatio(s1, s2, score_cutoff=0):
    if not s1 or not s2:                      # fuzzywuzzy compatibility: empty -> 0
        return 0
    return ratio(s1, s2, score_cutoff)


def _partial_ratio_impl(s1, s2, score_cutoff):
    """len(s1) <= len(s2); score_cutoff in [0, 1].  Best normalised Indel similarity of s1 against the prefixes of s2
    shorter than s1, every window of len(s1) characters, and the suffixes of s2 not longer than s1."""
    chars = set(s1)
    len1, len2 = len(s1), len(s2)
    best = 0.0
    for i in range(1, len1):
        if s2[i - 1] not in chars:
            continue
        r = _indel_norm_sim(s1, s2[:i], score_cutoff)
        if r > best:
            best = score_cutoff = r
            if best == 1:
                return 100
    for i in range(len2 - len1):
        if s2[i + len1 - 1] not in chars:
            continue
        r = _indel_norm_sim(s1, s2[i:i + len1], score_cutoff)
        if r > best:
            best = score_cutoff = r
            if best == 1:
                return 100
    for i in range(len2 - len1, len2):
        if s2[i] not in chars:
            continue
        r = _indel_norm_sim(s1, s2[i:], score_cutoff)
        if r > best:
            best = score_cutoff = r
            if best == 1:
                return 100
    return best * 100


def partial_ratio(s1, s2, score_cutoff=0):
    if not s1 and not s2:
        return 100.0
    shorter, longer = (s1, s2) if len(s1) <= len(s2) else (s2, s1)
    res = _partial_ratio_impl(shorter, longer, score_cutoff / 100)
    if res != 100 and len(s1) == len(s2):
        score_cutoff = max(score_cutoff, res)
        res2 = _partial_ratio_impl(longer, shorter, score_cutoff / 100)
        if res2 > res:
            res = res2
    return res if res >= score_cutoff else 0


def token_sort_ratio(s1, s2, score_cutoff=0):
    return ratio(_join(sorted(_split(s1))), _join(sorted(_split(s2))), score_cutoff)


def token_set_ratio(s1, s2, score_cutoff=0):
    if score_cutoff > 100:
        return 0
    tokens_a, tokens_b = set(_split(s1)), set(_split(s2))
    if not tokens_a or not tokens_b:
        return 0
    intersect = tokens_b & tokens_a
    diff_ab, diff_ba = tokens_a - tokens_b, tokens_b - tokens_a
    if intersect and (not diff_ab or not diff_ba):
        return 100
    ab, ba = _join(sorted(diff_ab)), _join(sorted(diff_ba))
    ab_len, ba_len = len(ab), len(ba)
    sect_len = len(_join(sorted(intersect)))
    sect_ab_len = sect_len + (sect_len != 0) + ab_len
    sect_ba_len = sect_len + (sect_len != 0) + ba_len
    result = 0.0
    cutoff_distance = ceil((sect_ab_len + sect_ba_len) * (1 - score_cutoff / 100))
    dist = indel_distance(ab, ba)
    if dist <= cutoff_distance:
        result = _norm_distance(dist, sect_ab_len + sect_ba_len, score_cutoff)
    if not sect_len:
        return result
    sect_ab_ratio = _norm_distance((sect_len != 0) + ab_len, sect_len + sect_ab_len, score_cutoff)
    sect_ba_ratio = _norm_distance((sect_len != 0) + ba_len, sect_len + sect_ba_len, score_cutoff)
    return max(result, sect_ab_ratio, sect_ba_ratio)


def token_ratio(s1, s2, score_cutoff=0):
    return max(token_set_ratio(s1, s2, score_cutoff), token_sort_ratio(s1, s2, score_cutoff))


def partial_token_sort_ratio(s1, s2, score_cutoff=0):
    return partial_ratio(_join(sorted(_split(s1))), _join(sorted(_split(s2))), score_cutoff)


def partial_token_set_ratio(s1, s2, score_cutoff=0):
    tokens_a, tokens_b = set(_split(s1)), set(_split(s2))
    if not tokens_a or not tokens_b:
        return 0
    if tokens_a & tokens_b:
        return 100
    return partial_ratio(_join(sorted(tokens_a - tokens_b)), _join(sorted(tokens_b - tokens_a)), score_cutoff)


def partial_token_ratio(s1, s2, score_cutoff=0):
    split_a, split_b = _split(s1), _split(s2)
    tokens_a, tokens_b = set(split_a), set(split_b)
    if tokens_a & tokens_b:
        return 100
    diff_ab, diff_ba = tokens_a - tokens_b, tokens_b - tokens_a
    result = partial_ratio(_join(sorted(split_a)), _join(sorted(split_b)), score_cutoff)
    if len(split_a) == len(diff_ab) and len(split_b) == len(diff_ba):
        return result
    score_cutoff = max(score_cutoff, result)
    return max(result, partial_ratio(_join(sorted(diff_ab)), _join(sorted(diff_ba)), score_cutoff))


def WRatio(s1, s2, score_cutoff=0):
    UNBASE_SCALE = 0.95
    if not s1 or not s2:
        return 0
    len1, len2 = len(s1), len(s2)
    len_ratio = len1 / len2 if len1 > len2 else len2 / len1
    end_ratio = ratio(s1, s2, score_cutoff)
    if len_ratio < 1.5:
        score_cutoff = max(score_cutoff, end_ratio) / UNBASE_SCALE
        return max(end_ratio, token_ratio(s1, s2, score_cutoff) * UNBASE_SCALE)
    PARTIAL_SCALE = 0.9 if len_ratio <= 8.0 else 0.6
    score_cutoff = max(score_cutoff, end_ratio) / PARTIAL_SCALE
    end_ratio = max(end_ratio, partial_ratio(s1, s2, score_cutoff) * PARTIAL_SCALE)
    score_cutoff = max(score_cutoff, end_ratio) / UNBASE_SCALE
    return max(end_ratio, partial_token_ratio(s1, s2, score_cutoff) * UNBASE_SCALE * PARTIAL_SCALE)


SCORERS = {"ratio": ratio, "QRatio": QRatio, "partial_ratio": partial_ratio, "token_sort_ratio": token_sort_ratio,
           "token_set_ratio": token_set_ratio, "token_ratio": token_ratio, "partial_token_sort_ratio": partial_token_sort_ratio,
           "partial_token_set_ratio": partial_token_set_ratio, "partial_token_ratio": partial_token_ratio, "WRatio": WRatio}


def extract_one(query, choices, scorer=WRatio, score_cutoff=0, exclude_index=None):
    """process.extractOne (processor=None): the first choice with the maximal score among those with score >= score_cutoff;
    -> (choice, score, index) or None.  exclude_index: skip that position (the matchers' self-match mode)."""
    best = None
    cutoff = score_cutoff
    for i, c in enumerate(choices):
        if i == exclude_index:
            continue
        s = scorer(query, c, cutoff)
        if s >= cutoff and (best is None or s > best[1]):
            best = (c, s, i)
            cutoff = s
            if s == 100:
                break
    return best
