"""Deterministic synthetic Arrow-columnar workloads.

The reference's demo_app reads a user file (apps/demo_app.cc:297-330); there is no dataset in this
environment, so the benchmark and the tests generate TPC-H-lineitem-like fixed-width Arrow value
buffers instead:  sorted int64 keys, int32 dictionary indices of 7 skewed categories, and float64
two-decimal prices, concatenated column-major in equal thirds.
"""
import os

import numpy as np

SEED = 20261018
_DICT_P = np.array([.30, .20, .15, .12, .10, .08, .05])


def col_sorted_int64(nbytes, rng):
    n = nbytes // 8
    return np.cumsum(rng.integers(0, 4, size=n, dtype=np.int64)).astype("<i8").view(np.uint8)


def col_dict_int32(nbytes, rng):
    n = nbytes // 4
    return rng.choice(7, size=n, p=_DICT_P).astype("<i4").view(np.uint8)


def col_price_f64(nbytes, rng):
    n = nbytes // 8
    return (rng.integers(90_000, 10_500_000, size=n, dtype=np.int64) / 100.0).astype("<f8").view(np.uint8)


COLUMNS = {"sorted_int64": col_sorted_int64, "dict_int32": col_dict_int32, "price_f64": col_price_f64}


def column(name, nbytes, seed=SEED):
    """nbytes of one column type (nbytes rounded down to the value width)."""
    return np.ascontiguousarray(COLUMNS[name](nbytes, np.random.default_rng(seed)))


def lineitem_like(nbytes, seed=SEED):
    """nbytes (multiple of 24 recommended) of the three columns in equal thirds, column-major."""
    rng = np.random.default_rng(seed)
    third = (nbytes // 3) // 8 * 8
    parts = [col_sorted_int64(third, rng), col_dict_int32(third, rng)]
    parts.append(col_price_f64(nbytes - 2 * third, rng))
    out = np.concatenate(parts)
    if out.size < nbytes:  # tail when nbytes - 2*third is not a multiple of 8
        out = np.concatenate([out, np.zeros(nbytes - out.size, np.uint8)])
    return np.ascontiguousarray(out)


def edge_cases(seg, seed=SEED):
    """BASELINE.json config 5 inputs: name -> uint8 array."""
    rng = np.random.default_rng(seed)
    cases = {}
    for n in (0, 1, 7, 8, seg - 1, seg, seg + 1, 3 * seg + 17):
        cases[f"random_{n}"] = np.frombuffer(rng.bytes(n), np.uint8).copy()
        cases[f"zeros_{n}"] = np.zeros(n, np.uint8)
    cases["ab"] = np.frombuffer(b"ab" * (seg // 2 + 5), np.uint8).copy()
    p258 = np.frombuffer(rng.bytes(258), np.uint8)
    cases["period258"] = np.tile(p258, (2 * seg) // 258 + 1).copy()
    p32k = np.frombuffer(rng.bytes(32768), np.uint8)
    cases["period32768"] = np.tile(p32k, 3).copy()
    p32k1 = np.frombuffer(rng.bytes(32769), np.uint8)
    cases["period32769"] = np.tile(p32k1, 3).copy()   # just beyond the 32 KiB window
    text = (b"the quick brown fox jumps over the lazy dog. " * 40 +
            b"pack my box with five dozen liquor jugs! " * 30)
    cases["text"] = np.frombuffer(text * (seg // len(text) + 2), np.uint8).copy()
    ramp = (np.arange(4 * seg, dtype=np.uint32) * 2654435761 >> 13).astype(np.uint8)
    cases["lowentropy"] = (ramp & 0x0F).astype(np.uint8)
    return cases


# ---- ratio corpus (round 2): inputs beyond the benchmark mix, for the "within 5 % of zlib level 1" gate ----
def _files_bytes(paths, nbytes):
    out = bytearray()
    for p in paths:
        try:
            with open(p, "rb") as f:
                out += f.read()
        except OSError:
            continue
        if len(out) >= nbytes:
            break
    return np.frombuffer(bytes(out[:nbytes]), np.uint8).copy()


def text_source(nbytes):
    """Program text: the Python standard library's own sources (present in this image), concatenated by name."""
    import glob
    import sysconfig
    return _files_bytes(sorted(glob.glob(os.path.join(sysconfig.get_paths()["stdlib"], "*.py"))), nbytes)


def elf_binary(nbytes):
    """Machine code + tables: the interpreter's ELF image."""
    import sys
    return _files_bytes([os.path.realpath(sys.executable)], nbytes)


def char_strings(nbytes, width=10, seed=SEED):
    """CHAR(width) column of 7 low-cardinality values (l_shipmode-like), space padded."""
    rng = np.random.default_rng(seed)
    vals = [b"AIR", b"FOB", b"MAIL", b"RAIL", b"REG AIR", b"SHIP", b"TRUCK"]
    table = np.frombuffer(b"".join(v.ljust(width) for v in vals), np.uint8).reshape(len(vals), width)
    idx = rng.integers(0, len(vals), size=nbytes // width + 1)
    return np.ascontiguousarray(table[idx].reshape(-1)[:nbytes])


def word_strings(nbytes, vocab=3000, seed=SEED):
    """Free text over a fixed vocabulary (l_comment-like): words of 2..10 lower-case letters, Zipf-ish use."""
    rng = np.random.default_rng(seed)
    words = [bytes(rng.integers(97, 123, size=int(rng.integers(2, 11)), dtype=np.uint8)) + b" " for _ in range(vocab)]
    p = 1.0 / np.arange(1, vocab + 1) ** 0.5
    picks = rng.choice(vocab, size=nbytes // 4 + 16, p=p / p.sum())
    out = b"".join(words[i] for i in picks)
    return np.frombuffer(out[:nbytes], np.uint8).copy()


def period_rows(nbytes, period=4096, seed=SEED):
    """One random row of `period` bytes repeated: redundancy only at distance `period`."""
    rng = np.random.default_rng(seed)
    row = np.frombuffer(rng.bytes(period), np.uint8)
    return np.ascontiguousarray(np.tile(row, nbytes // period + 1)[:nbytes])


def ratio_corpus(nbytes):
    """name -> uint8 array of (up to) nbytes: the inputs the compressed size is gated on against zlib level 1."""
    c = {"lineitem_mix": lineitem_like(nbytes)}
    for name in COLUMNS:
        c[name] = column(name, nbytes)
    c["text_source"] = text_source(nbytes)
    c["elf_binary"] = elf_binary(nbytes)
    c["char10_strings"] = char_strings(nbytes)
    c["word_strings"] = word_strings(nbytes)
    c["period4096_rows"] = period_rows(nbytes)
    return {k: v for k, v in c.items() if v.size}
