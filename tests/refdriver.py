"""Drive the UNMODIFIED reference functions on a ReadBatch (build container only).

Used by tests/golden/make_golden.py (fixture generation) and by the reference-pinning CPU
tests, which are skipped where /root/reference is not mounted (e.g. the GPU box).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402


def segments(batch):
    shim = ref_loader.shim()
    return [shim.AlignedSegment.from_sam_line(l) for l in batch.sam_lines()]


def ref_tables(L, primers, offset):
    """primers: sorted [(start, end)] -> the reference's two per-position lists (None = uncovered)."""
    ref = ref_loader.load_reference()
    return ref.find_overlapping_primers(L, [(int(s), int(e)) for s, e in primers], offset)


def ref_trim(batch, L, primers, offset=0, min_quality=20, window=4, min_length=30, include_no_primer=False):
    """AmpliPy.py:896-911 on every read.  Returns (segments, per-read dicts)."""
    ref = ref_loader.load_reference()
    prim = [(int(s), int(e)) for s, e in primers]
    mn, mx = ref.find_overlapping_primers(L, prim, offset)
    max_primer_len = max(e - s for s, e in prim)
    segs = segments(batch)
    out = []
    for s in segs:
        if s.is_unmapped or s.cigartuples is None:
            out.append({"skipped": True})
            continue
        ts, te, tq = ref.trim_read(s, mn, mx, max_primer_len, min_quality, window)
        keep = s.reference_length >= min_length and (ts or te or include_no_primer)
        out.append({"skipped": False, "pos": s.reference_start, "cigar": list(s.cigartuples), "ts": bool(ts),
                    "te": bool(te), "tq": bool(tq), "keep": bool(keep), "reference_length": s.reference_length})
    return segs, out


def ref_pileup(segs, L, min_quality=20):
    """AmpliPy.py:892, 902, 915.  Returns the reference's list of per-position dicts."""
    ref = ref_loader.load_reference()
    counts = [{'A': 0, 'C': 0, 'G': 0, 'T': 0, 'N': 0, '-': 0} for _ in range(L)]
    for s in segs:
        if s.is_unmapped or s.cigartuples is None:
            continue
        ref.update_base_counts(counts, s, min_quality)
    return counts


def ref_call(counts, ref_seq, min_depth_consensus=10, min_freq_consensus=0.0, min_depth_variants=1,
             min_freq_variants=0.03, unknown_symbol='N'):
    """AmpliPy.py:919-952 restated around the reference's own alleles_from_counts (the loop body is
    inline in run_amplipy and bound to pysam objects, so it is re-driven here line by line)."""
    ref = ref_loader.load_reference()
    L = len(counts)
    cons = [unknown_symbol] * L
    depth = []
    alleles = []
    variants = []
    for p in range(L):
        ref_symbol = ref_seq[p]
        total, al = ref.alleles_from_counts(counts[p])
        depth.append(total)
        alleles.append(al)
        if len(al) != 0 and al[0][0] >= min_depth_consensus and al[0][1] >= min_freq_consensus:
            cons[p] = al[0][2]
        tot_count = 0; rc = 0; rf = 0; asym = []; acnt = []; afreq = []
        for count, freq, symbol in al:
            tot_count += count
            if symbol == ref_symbol:
                rc = count; rf = freq
            elif freq >= min_freq_variants:
                asym.append(symbol); acnt.append(count); afreq.append(freq)
        if tot_count >= min_depth_variants and len(asym) != 0:
            if rc >= min_depth_variants and rf >= min_freq_variants:
                gt = tuple(range(len(asym) + 1))
            else:
                gt = tuple(range(1, len(asym) + 1))
            variants.append((p, ref_symbol, asym, total, rc, acnt, float(rf), afreq, gt))
    return {"depth": depth, "alleles": alleles, "consensus": "".join(cons), "variants": variants}


def counts_to_arrays(counts):
    """list-of-dicts -> (int64[6, L], {(pos, str): n}) in the oracle's layout."""
    import numpy as np
    L = len(counts)
    arr = np.zeros((6, L), np.int64)
    ins = {}
    for p, d in enumerate(counts):
        for k, v in d.items():
            if k in "ACGTN-" and len(k) == 1:
                arr["ACGTN-".index(k), p] = v
            else:
                ins[(p, k)] = v
    return arr, ins
