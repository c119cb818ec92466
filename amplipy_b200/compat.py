"""Function-level drop-ins with the reference's own signatures, for callers that drive the reference
read by read instead of through its command line:

    find_overlapping_primers(ref_genome_length, primers, primer_pos_offset)          AmpliPy.py:174-209
    trim_read(s, min_primer_start, max_primer_end, max_primer_len, min_quality, sliding_window_width)   426-687
    update_base_counts(symbol_counts_at_ref_pos, s, min_quality)                     690-753
    alleles_from_counts(symbol_counts)                                                756-771

``s`` is anything with pysam.AlignedSegment's attributes (reference_start, cigartuples, flag or is_paired / is_reverse /
is_unmapped, template_length, query_sequence, query_qualities); trim_read mutates its cigartuples / reference_start and
update_base_counts mutates the list of per-position dicts, exactly as the reference's functions do.  Every call is ONE batch of
one read through the CUDA path (amp_process_host / amp_call) -- correct but latency-bound; real workloads should use the batch
forms (engine.Engine.process / .call), of which these are the degenerate case.  There is no CPU fallback."""
import numpy as np

from . import primers as _primers
from .batch import ReadBatch
from .engine import F_TRIM_END, F_TRIM_QUAL, F_TRIM_START, Engine

_FIXED = "ACGTN-"
_engines = {}


def find_overlapping_primers(ref_genome_length, primers, primer_pos_offset):
    mn, mx = _primers.find_overlapping_primers(ref_genome_length, primers, primer_pos_offset)
    return [None if v < 0 else int(v) for v in mn], [None if v < 0 else int(v) for v in mx]


def _flag_of(s):
    if hasattr(s, "flag"):
        return int(s.flag)
    return (1 if s.is_paired else 0) | (4 if s.is_unmapped else 0) | (16 if s.is_reverse else 0)


def _batch_of(s):
    ops = [(int(op), int(n)) for op, n in (s.cigartuples or [])]
    seq = s.query_sequence or ""
    qual = list(s.query_qualities) if s.query_qualities is not None else [255] * len(seq)
    return ReadBatch.from_records([(int(s.reference_start), _flag_of(s), int(s.template_length), ops, seq.upper(), qual)])


def _engine(key, **kw):
    e = _engines.get(key)
    if e is None:
        if len(_engines) > 8:
            _engines.pop(next(iter(_engines))).close()
        e = _engines[key] = Engine(**kw)
    return e


def trim_read(s, min_primer_start, max_primer_end, max_primer_len, min_quality, sliding_window_width):
    """Returns (trimmed_primer_start, trimmed_primer_end, trimmed_quality) and rewrites s.cigartuples / s.reference_start."""
    L = len(min_primer_start)
    key = ("trim", id(min_primer_start), id(max_primer_end), L, max_primer_len, min_quality, sliding_window_width)
    mn = np.array([-1 if v is None else v for v in min_primer_start], np.int32)
    mx = np.array([-1 if v is None else v for v in max_primer_end], np.int32)
    eng = _engine(key, ref_len=L, primer_tables=(mn, mx), max_primer_len=max_primer_len, min_quality=min_quality,
                  sliding_window_width=sliding_window_width, min_length=1, include_no_primer=True)
    t = eng.process(_batch_of(s), trim=True, pileup=False)
    eng.raise_on_device_errors()
    fl = int(t.flags[0])
    s.cigartuples = t.cigartuples(0)
    s.reference_start = int(t.pos[0])
    return bool(fl & F_TRIM_START), bool(fl & F_TRIM_END), bool(fl & F_TRIM_QUAL)


def update_base_counts(symbol_counts_at_ref_pos, s, min_quality):
    L = len(symbol_counts_at_ref_pos)
    eng = _engine(("pile", L, min_quality), ref_len=L, min_quality=min_quality)
    eng.reset()
    eng.process(_batch_of(s), trim=False, pileup=True)
    eng.raise_on_device_errors()
    counts = eng.counts()
    for ch, p in zip(*np.nonzero(counts)):
        symbol_counts_at_ref_pos[int(p)][_FIXED[int(ch)]] += int(counts[ch, p])
    ins = eng.insertions()
    for k in range(ins.k):
        d = symbol_counts_at_ref_pos[int(ins.pos[k])]
        d[ins.strs[k]] = d.get(ins.strs[k], 0) + int(ins.count[k])


def alleles_from_counts(symbol_counts):
    """(total_coverage, [(count, frequency, symbol), ...]) sorted as the reference sorts them (descending count, frequency, symbol)."""
    eng = _engine(("call", 1), ref_len=1, min_quality=0)
    eng.reset()
    fixed = np.zeros((6, 1), np.int32)
    extra = []
    for sym, c in symbol_counts.items():
        if len(sym) == 1 and sym in _FIXED:
            fixed[_FIXED.index(sym), 0] = c
        elif c:
            extra.append((sym, int(c)))
    eng.upload_counts(fixed)
    if extra:
        off = np.zeros(len(extra) + 1, np.int64)
        np.cumsum([len(x[0]) for x in extra], out=off[1:])
        eng.merge_insertions(np.zeros(len(extra), np.int32), np.zeros(len(extra), np.int32), np.array([x[1] for x in extra], np.int32),
                             off, np.frombuffer("".join(x[0] for x in extra).encode("latin-1"), np.uint8))
    res = eng.call("N", 0, 0.0, 0, 0.0)
    ins = eng.insertions()
    total = int(res.depth[0])
    if total == 0:
        return 0, list()
    out = []
    for ch in range(6):
        if fixed[ch, 0]:
            out.append((int(res.fixed_rank[0, ch]), (int(fixed[ch, 0]), float(res.fixed_freq[0, ch]), _FIXED[ch])))
    for k in range(ins.k):
        out.append((int(res.ins_rank[k]), (int(ins.count[k]), float(res.ins_freq[k]), ins.strs[k])))
    out.sort(key=lambda x: x[0])
    return total, [x[1] for x in out]
