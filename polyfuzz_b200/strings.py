"""Host marshalling of Python string lists into the layout the C ABI takes (include/pfz.h):
UTF-32 code points in one uint32 blob + int64 offsets (n+1 entries).  Python's len() counts code
points, which is what both the reference's n-gram slicing (polyfuzz/models/_tfidf.py:132) and
rapidfuzz operate on."""
import numpy as np


def pack_utf32(strings):
    n = len(strings)
    offsets = np.zeros(n + 1, dtype=np.int64)
    if n:
        np.cumsum(np.fromiter(map(len, strings), dtype=np.int64, count=n), out=offsets[1:])
    try:
        raw = "".join(strings).encode("utf-32-le", "surrogatepass")
    except TypeError as e:                       # non-str element
        raise TypeError("all elements of the string list must be str") from e
    blob = np.frombuffer(raw, dtype=np.uint32).copy()       # writable (torch.from_numpy)
    if blob.size != offsets[-1]:
        raise ValueError("string list could not be packed as UTF-32")
    return blob, offsets


def ngram_slot_bounds(offsets, lo, hi):
    """Upper bound of n-gram occurrences per string (before cleaning): sum_n max(0, len-n+1);
    returns (slots int64[n], occ_ptr int64[n+1])."""
    lens = np.diff(offsets)
    slots = np.zeros_like(lens)
    for n in range(lo, hi + 1):
        slots += np.maximum(lens - n + 1, 0)
    occ = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(slots, out=occ[1:])
    return slots, occ
