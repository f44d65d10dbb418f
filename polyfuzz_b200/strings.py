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
    hp = _hostpack()
    if hp is not None and hasattr(hp, "slots") and offsets.dtype == np.int64 and offsets.flags.c_contiguous and len(offsets) > 0:
        n = len(offsets) - 1
        slots = np.empty(n, dtype=np.int64); occ = np.empty(n + 1, dtype=np.int64)
        hp.slots(offsets, slots, occ, int(lo), int(hi))           # one C pass
        return slots, occ
    lens = np.diff(offsets)
    slots = np.zeros_like(lens)
    for n in range(lo, hi + 1):
        slots += np.maximum(lens - n + 1, 0)
    occ = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(slots, out=occ[1:])
    return slots, occ


try:
    import pyarrow as _pa
    import pyarrow.compute as _pc
except Exception:                                        # pragma: no cover
    _pa = _pc = None

# id(list) -> (list, Arrow array) of the lists packed by the current call; lets the frame assembly reuse them
ARROW_CACHE = {}


def _hostpack():
    """The C packer (polyfuzz_b200/csrc/pfz_hostpack.c), or None if it has not been built."""
    global _HP
    if _HP is False:
        try:
            from . import _pfz_hostpack as hp
            _HP = hp
        except Exception:
            _HP = None
    return _HP


_HP = False
_BYTES_PER_STRING = 40.0     # running guess for the one-pass ASCII packer


def pack_strings(strings):
    """(blob, offsets, arrow_or_None): `blob` is uint8 (pure-ASCII list: bytes == code points, a zero-copy view of
    the Arrow data buffer, widened to uint32 on the device) or uint32 (UTF-32).  The Arrow array is built once and
    cached for the frame assembly; anything Arrow cannot take (lone surrogates, non-str) falls back to pack_utf32."""
    n = len(strings)
    hp = _hostpack()
    if hp is not None and n and isinstance(strings, (list, tuple)):
        if hasattr(hp, "fill_ascii"):
            # the common case in ONE pass: a pure-ASCII list that fits a guessed byte buffer (the last list's bytes per string + slack)
            global _BYTES_PER_STRING
            guess = int(n * _BYTES_PER_STRING) + 4096
            big = np.empty(guess, dtype=np.uint8)
            offsets = np.empty(n + 1, dtype=np.int64)
            total = hp.fill_ascii(strings, big, offsets)          # TypeError on non-str
            if total >= 0:
                _BYTES_PER_STRING = max(8.0, 1.25 * total / n)
                blob = big[:total] if 2 * total >= guess else big[:total].copy()
                return (blob if total else np.zeros(0, np.uint8)), offsets, None
            _BYTES_PER_STRING = min(256.0, 2.0 * _BYTES_PER_STRING)
        total, ascii_only = hp.scan(strings)                  # one C pass; TypeError on non-str
        blob = np.empty(max(total, 1), dtype=np.uint8 if ascii_only else np.uint32)
        offsets = np.empty(n + 1, dtype=np.int64)
        hp.fill(strings, blob, offsets, 1 if ascii_only else 4)
        return blob[:total], offsets, None
    if _pa is None or n == 0:
        b, o = pack_utf32(strings)
        return b, o, None
    try:
        arr = _pa.array(strings, type=_pa.large_string())
        if arr.null_count:
            raise TypeError("all elements of the string list must be str")
        lens = _pc.utf8_length(arr).to_numpy(zero_copy_only=False).astype(np.int64, copy=False)
    except TypeError:
        raise
    except Exception:
        b, o = pack_utf32(strings)
        return b, o, None
    ARROW_CACHE[id(strings)] = (strings, arr)
    bufs = arr.buffers()
    byte_off = np.frombuffer(bufs[1], dtype=np.int64, count=n + 1, offset=arr.offset * 8)
    total_bytes = int(byte_off[-1] - byte_off[0])
    total_cp = int(lens.sum())
    if total_bytes == total_cp:                           # pure ASCII
        data = np.frombuffer(bufs[2], dtype=np.uint8, count=total_bytes, offset=int(byte_off[0])) if total_bytes else np.zeros(0, np.uint8)
        offsets = byte_off - byte_off[0] if byte_off[0] else byte_off
        return data, np.ascontiguousarray(offsets), arr
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    blob = np.frombuffer("".join(strings).encode("utf-32-le", "surrogatepass"), dtype=np.uint32).copy()
    return blob, offsets, arr
