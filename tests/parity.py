"""Shared parity checks: the same assertions are applied to the CPU emulation of the kernels (CPU
tests) and to the CUDA engine through the C ABI (GPU tests)."""
import numpy as np

import golden_io
from amplipy_b200 import calling
from amplipy_b200.primers import find_overlapping_primers, max_primer_len


def engine_kwargs(meta):
    p = meta["params"]
    prim = [tuple(x) for x in meta["primers"]]
    tables = find_overlapping_primers(meta["L"], prim, p["offset"])
    return dict(ref_len=meta["L"], primer_tables=tables, max_primer_len=max_primer_len(prim),
                min_quality=p["min_quality"], sliding_window_width=p["window"], min_length=p["min_length"],
                include_no_primer=p["include_no_primer"])


def check_trim(t, arr, b):
    assert np.array_equal(t.flags, arr["t_flags"]), np.flatnonzero(t.flags != arr["t_flags"])[:10]
    assert np.array_equal(t.pos, arr["t_pos"])
    assert np.array_equal(t.ncig.astype(np.int32), arr["t_ncig"])
    # compare rows only up to ncig (slack beyond is unspecified)
    for i in np.flatnonzero((t.flags & 32) == 0):
        a = int(b.cig_off[i]) + 3 * int(i)
        n = int(t.ncig[i])
        assert np.array_equal(t.cigar[a:a + n], arr["t_cigar"][a:a + n]), (i, t.cigartuples(int(i)))


def check_case_aio(make_engine, name):
    """trim + pileup fused (the `aio` data flow), then calling; everything against the golden fixture."""
    b, meta, arr = golden_io.load_case(name)
    p = meta["params"]
    eng = make_engine(**engine_kwargs(meta))
    t = eng.process(b, trim=True, pileup=True)
    assert eng.error_flags() == 0
    check_trim(t, arr, b)
    counts = eng.counts()
    assert np.array_equal(counts, arr["counts_aio"]), np.argwhere(counts != arr["counts_aio"])[:10]
    ins = eng.insertions()
    assert ins.as_dict() == golden_io.ins_from_meta(meta)
    ref_seq = bytes(arr["ref_seq"]).decode()
    res = eng.call(ref_seq, p["min_depth_consensus"], p["min_freq_consensus"], p["min_depth_variants"],
                   p["min_freq_variants"])
    assert np.array_equal(res.depth.astype(np.int64), arr["depth_aio"])
    assert calling.consensus_string(res, ins, 0, p["unknown_symbol"]) == meta["consensus"]
    got = calling.variant_records(res, ins, ref_seq, counts)
    want = meta["variants"]
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert [g[0], g[1], g[2], g[3], g[4], g[5]] == w[:6], (g, w)
        assert g[6] == w[6] and g[7] == w[7], (g, w)          # float64 frequencies: bit-exact
        assert list(g[8]) == w[8], (g, w)
    # allele order of every position (fixed_rank / ins_rank) against the reference's sorted lists
    check_allele_order(res, ins, counts, arr, meta)
    return eng


def check_allele_order(res, ins, counts, arr, meta):
    L = meta["L"]
    al_off = arr["al_off"]
    by_pos = {}
    for k in range(ins.k):
        by_pos.setdefault(int(ins.pos[k]), []).append(k)
    sym = meta["al_sym"]
    for p in range(L):
        a, b = int(al_off[p]), int(al_off[p + 1])
        mine = []
        for ch in range(6):
            if counts[ch, p]:
                mine.append((int(res.fixed_rank[p, ch]), calling.FIXED_SYMS[ch], int(counts[ch, p]), float(res.fixed_freq[p, ch])))
        for k in by_pos.get(p, ()):
            mine.append((int(res.ins_rank[k]), ins.strs[k], int(ins.count[k]), float(res.ins_freq[k])))
        mine.sort()
        assert [m[0] for m in mine] == list(range(b - a)), (p, mine)
        assert [m[1] for m in mine] == sym[a:b], (p, mine, sym[a:b])
        assert [m[2] for m in mine] == arr["al_count"][a:b].tolist()
        assert [m[3] for m in mine] == arr["al_freq"][a:b].tolist()


def check_case_pileup_only(make_engine, name):
    """variants/consensus subcommands: pileup of the input alignments as they are (no trimming)."""
    b, meta, arr = golden_io.load_case(name)
    kw = engine_kwargs(meta)
    kw["primer_tables"] = None
    eng = make_engine(**kw)
    eng.process(b, trim=False, pileup=True)
    assert eng.error_flags() == 0
    assert np.array_equal(eng.counts(), arr["counts_raw"])
    assert eng.insertions().as_dict() == {(int(a), s): int(c) for a, s, c in meta["insertions_raw"]}


def check_case_trim_only(make_engine, name):
    b, meta, arr = golden_io.load_case(name)
    eng = make_engine(**engine_kwargs(meta))
    t = eng.process(b, trim=True, pileup=False)
    check_trim(t, arr, b)
    assert not eng.counts().any()
    return eng, t, b, meta, arr


def check_case_pipeline(make_engine, name):
    """config 1 flow: trim -> (kept reads only) -> variants, as three separate subcommand runs do."""
    eng, t, b, meta, arr = check_case_trim_only(make_engine, name)
    tb, sel = t.trimmed_batch(only_kept=True)
    kw = engine_kwargs(meta)
    kw["primer_tables"] = None
    eng2 = make_engine(**kw)
    eng2.process(tb, trim=False, pileup=True)
    assert np.array_equal(eng2.counts(), arr["counts_kept"])
    assert eng2.insertions().as_dict() == {(int(a), s): int(c) for a, s, c in meta["insertions_kept"]}


def check_against_oracle(make_engine, oracle, b, g, prim, offset=0, mq=20, w=4, ml=30, inc=False, ins_slots=0, arena=0):
    """trim + pileup + call of one engine (CUDA through the C ABI, or the CPU emulation of the kernels) against the oracle."""
    L = len(g)
    mn, mx = oracle.find_overlapping_primers(L, prim, offset)
    tables = find_overlapping_primers(L, prim, offset)
    assert np.array_equal(tables[0], mn) and np.array_equal(tables[1], mx)
    mpl = max_primer_len(prim)
    want = oracle.trim_batch(b, L, mn, mx, mpl, mq, w, ml, inc)
    wc, wins, nerr = oracle.pileup_batch(b, L, mq, trimmed=want)
    assert nerr == 0
    eng = make_engine(ref_len=L, primer_tables=tables, max_primer_len=mpl, min_quality=mq, sliding_window_width=w,
                      min_length=ml, include_no_primer=inc, ins_slots=ins_slots, ins_arena_bytes=arena)
    t = eng.process(b, trim=True, pileup=True)
    assert eng.error_flags() == 0
    assert np.array_equal(t.flags, want["flags"])
    assert np.array_equal(t.pos, want["pos"])
    assert np.array_equal(t.ncig.astype(np.int32), want["ncig"])
    # rows up to ncig: build a mask of live output words
    row0 = b.cig_off[:-1].astype(np.int64) + 3 * np.arange(b.n, dtype=np.int64)
    live = np.repeat(row0, want["ncig"]) + (np.arange(int(want["ncig"].sum())) - np.repeat(np.cumsum(want["ncig"]) - want["ncig"], want["ncig"]))
    assert np.array_equal(t.cigar[live], want["cigar"][live])
    counts = eng.counts()
    assert np.array_equal(counts.astype(np.int64), wc)
    ins = eng.insertions()
    assert ins.as_dict() == wins
    res = eng.call(g)
    ores = oracle.call(wc, wins, g)
    assert np.array_equal(res.depth.astype(np.int64), ores["depth"])
    assert calling.consensus_string(res, ins) == oracle.consensus_string(ores)
    got = calling.variant_records(res, ins, g, counts)
    want_v = oracle.variant_records(ores, g)
    assert len(got) == len(want_v)
    for a, c in zip(got, want_v):
        assert tuple(a[:6]) == tuple(c[:6]) and a[6] == c[6] and a[7] == c[7] and tuple(a[8]) == tuple(c[8]), (a, c)
    return eng, t
