"""CPU, build container only: the C oracle (oracle/amplipy_oracle.c) against the UNMODIFIED reference
(/root/reference/AmpliPy.py imported with the pysam shim) on freshly generated inputs -- trim, pileup and calling.
The committed fixtures under tests/golden/ pin the oracle on fixed inputs; this test pins it live on other seeds.
Skipped where the reference is not mounted (the GPU box)."""
import numpy as np
import pytest

from amplipy_b200 import synth
from amplipy_b200.batch import ReadBatch
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference is not mounted")


def _scheme(L, n_amp, seed, n_alt=0):
    g = synth.random_genome(L, 7 + seed)
    primers, amps = synth.make_scheme(L, n_amp, seed=seed, n_alt=n_alt)
    return g, sorted((int(s), int(e)) for s, e, _ in primers), amps


def _compare(oracle, b, g, prim, offset=0, mq=20, w=4, ml=30, inc=False, mdc=10, mfc=0.0, mdv=1, mfv=0.03):
    import refdriver
    L = len(g)
    segs, ref_out = refdriver.ref_trim(b, L, prim, offset, mq, w, ml, inc)
    mn, mx = oracle.find_overlapping_primers(L, prim, offset)
    rmn, rmx = refdriver.ref_tables(L, prim, offset)
    assert [(-1 if v is None else v) for v in rmn] == mn.tolist() and [(-1 if v is None else v) for v in rmx] == mx.tolist()
    t = oracle.trim_batch(b, L, mn, mx, max(e - s for s, e in prim), mq, w, ml, inc)
    for i, r in enumerate(ref_out):
        if r["skipped"]:
            assert t["flags"][i] == oracle.F_SKIPPED, i
            continue
        want_flags = (1 if r["ts"] else 0) | (2 if r["te"] else 0) | (4 if r["tq"] else 0) | (8 if r["keep"] else 0)
        assert int(t["flags"][i]) == want_flags, (i, b.record(i)[:4], r, int(t["flags"][i]))
        assert int(t["pos"][i]) == r["pos"], (i, b.record(i)[:4], r)
        assert oracle.trimmed_cigartuples(b, t, i) == [tuple(x) for x in r["cigar"]], (i, b.record(i)[:4], r)
    rc = refdriver.ref_pileup(segs, L, mq)
    want_counts, want_ins = refdriver.counts_to_arrays(rc)
    counts, ins, nerr = oracle.pileup_batch(b, L, mq, trimmed=t)
    assert nerr == 0
    assert np.array_equal(counts, want_counts), np.argwhere(counts != want_counts)[:10]
    assert ins == want_ins
    rr = refdriver.ref_call(rc, g, mdc, mfc, mdv, mfv)
    orr = oracle.call(counts, ins, g, True, mdc, mfc, True, mdv, mfv)
    assert orr["depth"].tolist() == rr["depth"]
    assert oracle.consensus_string(orr) == rr["consensus"]
    got = oracle.variant_records(orr, g)
    assert len(got) == len(rr["variants"])
    for a, c in zip(got, rr["variants"]):
        assert (a[0], a[1], a[2], a[3], a[4], a[5]) == (c[0], c[1], c[2], c[3], c[4], c[5]), (a, c)
        assert a[6] == c[6] and a[7] == c[7] and tuple(a[8]) == tuple(c[8]), (a, c)      # float64 frequencies bit-exact
    # allele order of every position
    for p in range(L):
        a0, a1 = int(orr["al_off"][p]), int(orr["al_off"][p + 1])
        mine = [(int(orr["al_count"][k]), float(orr["al_freq"][k]), orr["sym"](int(orr["al_sym"][k]))) for k in range(a0, a1)]
        assert mine == [tuple(x) for x in rr["alleles"][p]], p


@pytest.mark.parametrize("seed,offset,mq,w,ont", [(101, 0, 20, 4, False), (102, 2, 20, 4, True), (103, 5, 0, 1, False),
                                                 (104, 10, 30, 10, True), (105, 0, 11, 4, False), (106, 3, 20, 6, True)])
def test_fuzz_reads(oracle_lib, seed, offset, mq, w, ont):
    Lf = 900
    g, prim, _ = _scheme(Lf, 6, seed % 4 + 1)
    recs = synth.fuzz_records(Lf, 700, seed=seed, ont_like=ont)
    b = ReadBatch.from_records(sorted(recs, key=lambda r: r[0]))
    _compare(oracle_lib, b, g, prim, offset=offset, mq=mq, w=w, inc=bool(seed & 1))


@pytest.mark.parametrize("seed", [111, 112, 113])
def test_reads_at_the_ends_of_the_reference(oracle_lib, seed):
    """Reads at position 0 and reads ending on the last base, with a primer that reaches the last base: a read swallowed
    by the start clip ends up at pos == L with an all-S CIGAR, which the reference piles up without touching anything."""
    L = 240
    g = synth.random_genome(L, seed)
    prim = [(0, 22), (60, 84), (150, 171), (L - 24, L)]
    recs = synth.fuzz_records(L, 500, seed=seed, max_len=40, edges=0.5)
    recs += [(L - 2, 0, 0, [(0, 2)], "AC", [30, 30]), (L - 6, 16, 0, [(0, 6)], "ACGTAC", [30] * 6), (0, 0, 0, [(0, 12)], "ACGTACGTACGT", [35] * 12)]
    b = ReadBatch.from_records(sorted(recs, key=lambda r: r[0]))
    _compare(oracle_lib, b, g, prim, offset=seed % 3, ml=5)


def test_ont_like(oracle_lib):
    L = 3000
    g, prim, amps = _scheme(L, 9, 3)
    b = synth.ont_batch(g, amps, 500, seed=121)
    _compare(oracle_lib, b, g, prim, mq=10)
    _compare(oracle_lib, b, g, prim, mq=20, mfv=0.2)


def test_illumina_with_indels(oracle_lib):
    L = 4000
    g, prim, amps = _scheme(L, 12, 2, n_alt=3)
    b = synth.illumina_batch(g, amps, 3000, seed=122, p_ins=0.2, p_del=0.2, p_clip=0.2, p_hard=0.05,
                             snvs=[(700, "T", 0.4), (2500, "A", 0.05)])
    _compare(oracle_lib, b, g, prim)
    _compare(oracle_lib, b, g, prim, offset=4, mq=30, inc=True)
