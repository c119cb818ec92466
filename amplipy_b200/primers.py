"""Primer BED loading and the per-position primer tables (host side, tiny one-off work).

Mirrors AmpliPy.py:235-258 (``load_primers``) and 174-209 (``find_overlapping_primers``); the tables
are what the trim kernel looks up per read (AmpliPy.py:450-451)."""
import os

import numpy as np

ERROR_TEXT_EMPTY_BED = "Empty BED file"
ERROR_TEXT_FILE_NOT_FOUND = "File not found"
ERROR_TEXT_INVALID_BED_LINE = "Invalid primer BED line"


class InputError(Exception):
    """Carries the reference's error text; the CLI prints it the way AmpliPy.py:85-90 does."""


def load_primers(primer_fn):
    """4-column BED -> sorted [(start, end)] (AmpliPy.py:235-258).  Any other column count is an
    "Invalid primer BED line", as in the reference."""
    if not os.path.isfile(primer_fn):
        raise InputError("%s: %s" % (ERROR_TEXT_FILE_NOT_FOUND, primer_fn))
    with open(primer_fn, "r") as f:
        lines = f.read().strip().splitlines()
    primers = []
    for l in lines:
        parts = l.split("\t")
        try:
            if len(parts) != 4:
                raise ValueError
            primers.append((int(parts[1]), int(parts[2])))
        except ValueError:
            raise InputError("%s: %s" % (ERROR_TEXT_INVALID_BED_LINE, l))
    if len(primers) == 0:
        raise InputError("%s: %s" % (ERROR_TEXT_EMPTY_BED, primer_fn))
    primers.sort()
    return primers


def find_overlapping_primers(ref_genome_length, primers, primer_pos_offset):
    """For every reference position p: over primers with start - offset <= p < end + offset,
    min(start) and max(end); -1 where no primer covers p (None in AmpliPy.py:190-191).
    Same result as the reference's sweep (174-209), computed primer by primer on numpy slices."""
    L = int(ref_genome_length)
    mn = np.full(L, np.iinfo(np.int32).max, np.int32)
    mx = np.full(L, -1, np.int32)
    off = int(primer_pos_offset)
    for s, e in primers:
        lo, hi = max(s - off, 0), min(e + off, L)
        if lo < hi:
            np.minimum(mn[lo:hi], s, out=mn[lo:hi])
            np.maximum(mx[lo:hi], e, out=mx[lo:hi])
    mn[mx < 0] = -1
    return mn, mx


def max_primer_len(primers):
    return max(e - s for s, e in primers)   # AmpliPy.py:876
