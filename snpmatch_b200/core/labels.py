"""Vectorised handling of per-marker string labels (chromosome names, GT strings) on the host.

The reference maps names per marker (pandas `str.replace` on every element, parsers.py:161; a Python list of N labels,
pygwas/genotype.py:156-161).  Here a label column is first reduced to integer codes plus its few distinct values, and every
string operation runs on the distinct values only."""
import numpy as np
import pandas as pd


def factorize(labels):
    """(codes int64[n], uniques str[k]) with uniques in first-appearance order.  Labels that come in long runs (a file sorted
    by chromosome) are handled by run detection, anything else by pandas' hash-based factorize."""
    labels = np.asarray(labels).ravel()
    if labels.dtype.kind not in "US":
        labels = labels.astype("U")
    n = len(labels)
    if n == 0:
        return np.zeros(0, dtype=np.int64), labels.astype("U")
    # neighbours compared as integer words (a fixed-width unicode array is 4 bytes per character): several times faster than
    # NumPy's string comparison
    if labels.dtype.kind == "U" and labels.dtype.itemsize >= 4:
        wide = labels.dtype.itemsize % 8 == 0                      # an even number of characters: compare two at a time
        words = np.ascontiguousarray(labels).view(np.uint64 if wide else np.uint32).reshape(n, -1)
        differ = words[1:, 0] != words[:-1, 0]
        for k in range(1, words.shape[1]):
            differ |= words[1:, k] != words[:-1, k]
        change = np.flatnonzero(differ) + 1
    else:
        change = np.flatnonzero(labels[1:] != labels[:-1]) + 1
    if len(change) < max(64, n // 16):
        starts = np.concatenate([[0], change])
        run_codes, uniq = pd.factorize(labels[starts])
        return np.repeat(run_codes.astype(np.int64), np.diff(np.concatenate([starts, [n]]))), np.asarray(uniq).astype("U")
    codes, uniq = pd.factorize(labels)
    return codes.astype(np.int64), np.asarray(uniq).astype("U")


def map_labels(labels, fn, per_label=True):
    """fn applied to every label, computed on the distinct labels only: returns (mapped str[n], codes, mapped uniques); the
    per-marker strings are left out (None) with per_label=False."""
    codes, uniq = factorize(labels)
    mapped = np.array([fn(str(u)) for u in uniq], dtype="str") if len(uniq) else np.zeros(0, dtype="U1")
    if not per_label:
        return None, codes, mapped
    return (mapped[codes] if len(codes) else np.zeros(0, dtype="U1")), codes, mapped
