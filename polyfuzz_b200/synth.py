"""Seeded synthetic string lists shaped like the reference's benchmark inputs (the GPU box has no
copy of /root/reference/data, and no network):

  company_names(n)   -- BASELINE config 2 stand-in: company-name-like strings whose char-trigram
                        statistics are calibrated to data/company_names.json (100 000 names:
                        V = 13 264 trigrams, nnz = 1 310 412, P = sum_t df_t^2 = 3.54e9; heavy suffix
                        tokens such as 'inc'/'llc' in ~20 % of rows)  [SURVEY.md section 8d].
                        Measured for the default seed: V = 11 007, nnz = 1 360 993, P = 3.63e9,
                        272 terms carry 94 % of P (real: 306).
  uniform_strings(n) -- BASELINE config 5: lengths U[8,32], characters iid over [a-z0-9 ]
  titles(n)          -- BASELINE config 3 stand-in: movie-title-like strings (mixed case, punctuation,
                        a few non-ASCII letters), mean length ~16-18 code points, max < 200
"""
import numpy as np

_SUFFIX = [("LLC", 0.195), ("INC", 0.194), ("LP", 0.10), ("FUND", 0.055), ("TRUST", 0.045), ("LTD", 0.046),
           ("CORP", 0.046), ("CO", 0.02), ("", 0.299)]
_COMMON = ["FUND", "CAPITAL", "TRUST", "PARTNERS", "OF", "SERIES", "HOLDINGS", "ADVISORS", "II", "AMERICAN",
           "MANAGEMENT", "DISCIPLINED", "GROUP", "III", "INVESTMENT", "JOHN", "INVESTORS", "SECURITIES", "EQUITY",
           "ROBERT", "SERVICES", "INTERNATIONAL", "GLOBAL", "ENERGY", "FINANCIAL", "VENTURES", "ASSET", "THE",
           "TECHNOLOGIES", "OPPORTUNITY", "REAL", "ESTATE", "GROWTH", "INCOME", "MASTER", "PORTFOLIO", "RESOURCES",
           "WILLIAM", "JAMES", "DAVID", "MICHAEL", "ASSOCIATES", "BANK", "SYSTEMS", "HEALTH", "MEDICAL", "PROPERTIES",
           "ACQUISITION", "STRATEGIC", "VALUE", "INSTITUTIONAL", "OFFSHORE", "CREDIT", "INSURANCE", "LIFE", "NATIONAL",
           "FIRST", "NEW", "AMERICA", "PACIFIC", "ATLANTIC", "BIO", "PHARMA", "THERAPEUTICS", "SOLUTIONS", "MEDIA",
           "NETWORK", "DIGITAL", "DATA", "POWER", "OIL", "GAS", "MINING", "GOLD", "SILVER", "REALTY", "DEVELOPMENT"]
_LETTERS = "etaoinshrdlcumwfgypbvkjxqz"
_LETTER_P = np.array([12.0, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4, 2.2, 2.0, 2.0, 1.9, 1.5,
                      1.0, 0.8, 0.25, 0.2, 0.15, 0.1])
_VOWELS = set("aeiouy")


def _word_pool(rng, n_words, flat=0.42):
    """Pronounceable-ish pseudo words: letters from a flattened English unigram distribution with a
    vowel/consonant alternation bias, lengths 2..12."""
    p = _LETTER_P ** flat
    p = p / p.sum()
    is_v = np.array([c in _VOWELS for c in _LETTERS])
    pv = np.where(is_v, p, 0); pv /= pv.sum()
    pc = np.where(~is_v, p, 0); pc /= pc.sum()
    lens = rng.choice(np.arange(2, 13), size=n_words * 2, p=np.array([3, 8, 12, 14, 15, 14, 12, 9, 6, 4, 3]) / 100.0)
    words, out = set(), []
    for L in lens:
        cs = []
        prev_v = rng.random() < 0.35
        for _ in range(L):
            r = rng.random()
            want_v = (not prev_v) if r < 0.78 else prev_v
            c = _LETTERS[rng.choice(26, p=pv if want_v else pc)]
            cs.append(c); prev_v = c in _VOWELS
        w = "".join(cs)
        if w not in words:
            words.add(w); out.append(w.upper())
            if len(out) == n_words:
                break
    return out


def company_names(n=100_000, seed=0, common_frac=0.40, flat=0.42, words_p=(0.10, 0.34, 0.30, 0.17, 0.09), tail_zipf=0.9):
    rng = np.random.default_rng(seed)
    pool = _word_pool(rng, 30_000, flat)
    # Zipf-like tail over the generated pool, a fixed head of common business words
    tail_p = 1.0 / np.arange(1, len(pool) + 1) ** tail_zipf
    tail_p /= tail_p.sum()
    common_p = 1.0 / np.arange(1, len(_COMMON) + 1) ** 0.55
    common_p /= common_p.sum()
    suffix_p = np.array([p for _, p in _SUFFIX]); suffix_p /= suffix_p.sum()
    n_words = rng.choice([1, 2, 3, 4, 5], size=n, p=list(words_p))
    is_common = rng.random((n, 5)) < common_frac
    tail_ix = rng.choice(len(pool), size=(n, 5), p=tail_p)
    comm_ix = rng.choice(len(_COMMON), size=(n, 5), p=common_p)
    suf_ix = rng.choice(len(_SUFFIX), size=n, p=suffix_p)
    punct = rng.random(n)
    initials = rng.random((n, 2)) < 0.06
    letters = rng.integers(0, 26, size=(n, 2))
    out = []
    for i in range(n):
        ws = [(_COMMON[comm_ix[i, j]] if is_common[i, j] else pool[tail_ix[i, j]]) for j in range(n_words[i])]
        for j in range(2):
            if initials[i, j]:
                ws.insert(min(j, len(ws)), chr(65 + letters[i, j]))
        suf = _SUFFIX[suf_ix[i]][0]
        s = " ".join(ws)
        if suf:
            if punct[i] < 0.25:
                s += ", " + suf + "."
            elif punct[i] < 0.35 and len(suf) <= 3:
                s += " " + ".".join(suf) + "."
            else:
                s += " " + suf
        out.append(s)
    return out


_ALNUM_SPACE = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz0123456789 ", dtype=np.uint8)


def uniform_strings(n=1_000_000, seed=0, lo=8, hi=32):
    """lengths rng.integers(8,33), characters iid uniform over 37 symbols."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(lo, hi + 1, n)
    offs = np.zeros(n + 1, dtype=np.int64); np.cumsum(lens, out=offs[1:])
    chars = _ALNUM_SPACE[rng.integers(0, 37, int(offs[-1]))].tobytes().decode("ascii")
    return [chars[offs[i]:offs[i + 1]] for i in range(n)]


_TITLE_WORDS = ["The", "of", "and", "a", "in", "to", "Love", "Man", "Night", "My", "La", "Last", "Day", "Life", "Story",
                "Girl", "Le", "Der", "Die", "El", "de", "House", "Time", "World", "Dead", "Black", "Little", "Big", "King",
                "Lady", "Home", "One", "Two", "Secret", "Return", "Blood", "City", "War", "Dark", "Wild", "Christmas"]
_NONASCII = "éèüöäñçøåßàíóúâêô"


def titles(n, seed=0):
    rng = np.random.default_rng(seed)
    pool = [w.capitalize() for w in _word_pool(rng, 6000)]
    p = 1.0 / np.arange(1, len(pool) + 1) ** 0.7; p /= p.sum()
    out = []
    nw = rng.choice([1, 2, 3, 4, 5, 6, 8, 12], size=n, p=[0.18, 0.27, 0.24, 0.15, 0.08, 0.05, 0.02, 0.01])
    for i in range(n):
        ws = []
        for _ in range(nw[i]):
            if rng.random() < 0.35:
                ws.append(_TITLE_WORDS[rng.integers(len(_TITLE_WORDS))])
            else:
                w = pool[rng.choice(len(pool), p=p)]
                if rng.random() < 0.03:
                    k = rng.integers(len(w)); w = w[:k] + _NONASCII[rng.integers(len(_NONASCII))] + w[k + 1:]
                ws.append(w)
        s = " ".join(ws)
        r = rng.random()
        if r < 0.05:
            s += ": " + pool[rng.choice(len(pool), p=p)]
        elif r < 0.08:
            s += " " + str(rng.integers(2, 10))
        elif r < 0.10:
            s += "!"
        out.append(s)
    return out
