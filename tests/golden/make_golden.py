#!/usr/bin/env python3
"""Generate the committed golden fixtures by executing the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every expected value stored in tests/golden/*.npz is produced by the reference's own functions
(`trim_read`, `update_base_counts`, `alleles_from_counts`, `find_overlapping_primers` --
/root/reference/AmpliPy.py:174-209, 426-771) driven through oracle/pysam_shim; nothing in the
fixtures comes from this repo's oracle or CUDA code.  The inputs are seeded synthetics
(amplipy_b200/synth.py), the two example reads shipped with the reference, and the quirk vectors
of SURVEY.md section 8a-Q.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_io  # noqa: E402
import refdriver  # noqa: E402
from amplipy_b200 import synth  # noqa: E402
from amplipy_b200.batch import ReadBatch, cigar_string, parse_cigar  # noqa: E402
from oracle import ref_loader  # noqa: E402

EX = "/root/reference/example"
DEFAULTS = dict(offset=0, min_quality=20, window=4, min_length=30, include_no_primer=False,
                min_depth_consensus=10, min_freq_consensus=0.0, min_depth_variants=1, min_freq_variants=0.03,
                unknown_symbol="N")


def encode_trim(batch, rt):
    n = batch.n
    pos = batch.pos.copy()
    ncig = np.zeros(n, np.int32)
    flags = np.zeros(n, np.uint8)
    cig = np.zeros(int(batch.cig_off[-1]) + 3 * n, np.uint32)
    for i, r in enumerate(rt):
        a = int(batch.cig_off[i]) + 3 * i
        if r["skipped"]:
            flags[i] = 16
            ops = batch.cigartuples(i)
        else:
            pos[i] = r["pos"]
            flags[i] = (1 if r["ts"] else 0) | (2 if r["te"] else 0) | (4 if r["tq"] else 0) | (8 if r["keep"] else 0)
            ops = r["cigar"]
        ncig[i] = len(ops)
        assert len(ops) <= int(batch.cig_off[i + 1] - batch.cig_off[i]) + 3
        for k, (op, ln) in enumerate(ops):
            cig[a + k] = (ln << 4) | op
    return {"t_pos": pos, "t_ncig": ncig, "t_cigar": cig, "t_flags": flags}


def make_case(name, batch, L, ref_seq, primers, **kw):
    p = dict(DEFAULTS)
    p.update(kw)
    prim = sorted((int(s), int(e)) for s, e in primers)
    segs, rt = refdriver.ref_trim(batch, L, prim, p["offset"], p["min_quality"], p["window"], p["min_length"],
                                  p["include_no_primer"])
    arrays = encode_trim(batch, rt)
    # aio flow: every mapped read is piled up after trimming (AmpliPy.py:914-915)
    counts = refdriver.ref_pileup(segs, L, p["min_quality"])
    arr, ins = refdriver.counts_to_arrays(counts)
    arrays["counts_aio"] = arr.astype(np.int32)
    # trim -> variants pipeline: only reads that pass the write gate are seen by the second step
    kept = [s for s, r in zip(segs, rt) if not r["skipped"] and r["keep"]]
    ck = refdriver.ref_pileup(kept, L, p["min_quality"])
    arrk, insk = refdriver.counts_to_arrays(ck)
    arrays["counts_kept"] = arrk.astype(np.int32)
    # variants/consensus on the UNTRIMMED reads (pileup alone)
    raw = refdriver.ref_pileup(refdriver.segments(batch), L, p["min_quality"])
    arrr, insr = refdriver.counts_to_arrays(raw)
    arrays["counts_raw"] = arrr.astype(np.int32)
    call = refdriver.ref_call(counts, ref_seq, p["min_depth_consensus"], p["min_freq_consensus"],
                              p["min_depth_variants"], p["min_freq_variants"], p["unknown_symbol"])
    arrays["depth_aio"] = np.array(call["depth"], np.int64)
    mn, mx = refdriver.ref_tables(L, prim, p["offset"])
    arrays["min_primer_start"] = np.array([-1 if x is None else x for x in mn], np.int32)
    arrays["max_primer_end"] = np.array([-1 if x is None else x for x in mx], np.int32)
    arrays["ref_seq"] = np.frombuffer(ref_seq.encode(), np.uint8)
    al_off = np.zeros(L + 1, np.int64)
    al_count, al_freq, al_sym = [], [], []
    for q in range(L):
        for c, f, s in call["alleles"][q]:
            al_count.append(c); al_freq.append(f); al_sym.append(s)
        al_off[q + 1] = len(al_count)
    arrays["al_off"] = al_off
    arrays["al_count"] = np.array(al_count, np.int64)
    arrays["al_freq"] = np.array(al_freq, np.float64)
    meta = {"name": name, "L": L, "params": p, "primers": prim,
            "insertions": [[k[0], k[1], v] for k, v in sorted(ins.items())],
            "insertions_kept": [[k[0], k[1], v] for k, v in sorted(insk.items())],
            "insertions_raw": [[k[0], k[1], v] for k, v in sorted(insr.items())],
            "al_sym": al_sym, "consensus": call["consensus"],
            "variants": [[v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], list(v[8])] for v in call["variants"]]}
    golden_io.save_case(name, batch, meta, arrays)
    nt = sum(1 for r in rt if not r["skipped"] and (r["ts"] or r["te"] or r["tq"]))
    print("%-18s reads=%d trimmed=%d kept=%d ins_alleles=%d variants=%d" %
          (name, batch.n, nt, len(kept), len(ins), len(call["variants"])))
    return rt, counts


def rec(pos, cigar, flag=0, tlen=0, seq=None, qual=None):
    ops = parse_cigar(cigar)
    ql = sum(n for op, n in ops if op in (0, 1, 4, 7, 8))
    if seq is None:
        seq = ("ACGT" * (ql // 4 + 1))[:ql]
    if qual is None:
        qual = [40] * ql
    elif isinstance(qual, str):
        qual = [ord(c) - 33 for c in qual]
    return (pos, flag, tlen, ops, seq, qual)


def quirk_case():
    L = 200
    ref_seq = synth.random_genome(L, 99)
    primers = [(10, 30), (100, 120)]
    I, H = "I", "#"
    recs = [
        rec(10, "5H50M"), rec(12, "10M3D40M"), rec(12, "17M3D33M"), rec(12, "18M2I30M"),
        rec(12, "95M"), rec(12, "95M", flag=99, tlen=400), rec(12, "95M", flag=147, tlen=-400), rec(12, "10M"),
        rec(40, "50M", flag=16, qual=H * 10 + I * 40), rec(40, "50M", flag=16, qual=H + I * 49),
        rec(40, "50M", qual=I * 40 + H * 10), rec(40, "38M2D12M", qual=I * 40 + H * 10), rec(40, "50M", qual=H * 50),
        # pileup quirks (SURVEY.md 8a-Q second table), placed away from primers at pos 130+
        rec(130, "3M2I2D3M", seq="ACGTTACG"), rec(130, "3M3I3M", seq="ACGTTTACG", qual=[40, 40, 40, 40, 2, 40, 40, 40, 40]),
        rec(130, "3M2I3M", seq="ACGTTACG", qual=[40, 40, 40, 2, 40, 40, 40, 40]), rec(130, "3M2I2S", seq="ACGTTAC"),
        rec(130, "2I3M", seq="TTACG"), rec(130, "2S2I3M", seq="AGTTACG"), rec(0, "1S2I3M", seq="GTTACG"),
        rec(130, "2M2D2M", seq="ACGT", qual=[2, 2, 2, 2]), rec(130, "2M3N2M", seq="ACGT"),
        rec(130, "2I2D3M", seq="TTACG"), rec(150, "4M1I1D1I4M", seq="ACGTAGACGT"),
        rec(150, "4M2I4M4S", seq="ACGTAAACGTNNNN", qual=[40] * 10 + [2, 40, 2, 40]),
        (5, 4, 0, [], "ACGT", [30] * 4),               # unmapped
        (60, 0, 0, [], "ACGT", [30] * 4),              # mapped flag but no CIGAR
    ]
    b = ReadBatch.from_records(recs)
    rt, counts = make_case("quirks", b, L, ref_seq, primers, min_length=1)
    # cross-check a few entries against SURVEY.md's table (same reference, independent shim)
    expect = [(31, "21S29M"), (31, "16S34M"), (32, "17S33M"), (31, "21S29M"), (31, "19S69M7S"), (31, "19S76M"),
              (12, "88M7S"), (22, "10S"), (40, "11S39M"), (40, "50M"), (40, "39M11S"), (40, "38M2D1M11S"), (40, "50S")]
    for i, (p, c) in enumerate(expect):
        assert rt[i]["pos"] == p and cigar_string(rt[i]["cigar"]) == c, (i, rt[i], p, c)
    assert counts[132].get("GTTACG") == 1 and counts[135].get("GT") == 1 and counts[129].get("") == 1
    assert counts[0].get("TTA") == 1


def example_case():
    ref = ref_loader.load_reference()
    _, ref_seq = ref.load_ref_genome(os.path.join(EX, "example_reference.fas"))
    primers = ref.load_primers(os.path.join(EX, "example_primers.bed"))
    L = len(ref_seq)
    lines = []
    for fn in ("example_primer_trim_start.sam", "example_primer_trim_end.sam"):
        lines += [l for l in open(os.path.join(EX, fn)) if not l.startswith("@")]
    ex = ReadBatch.from_sam_lines(lines)
    # stand-in for the missing example_untrimmed_sorted.bam: seeded reads over amplicons snapped to example primers
    _, amps = synth.make_scheme(L, 98, seed=1)
    st = np.array(sorted(set(p[0] for p in primers)))
    en = np.array(sorted(set(p[1] for p in primers)))
    for a in amps:
        a[0] = st[np.argmin(np.abs(st - a[0]))]
        cand = en[en >= a[0] + 150]
        a[1] = cand[np.argmin(np.abs(cand - a[1]))]
    syn = synth.illumina_batch(ref_seq, amps, 4000, seed=1, snvs=[(241, "T", 1.0), (3037, "T", 0.6), (14408, "T", 0.3),
                                                                 (23403, "G", 0.95), (28881, "A", 0.08)])
    b = ReadBatch.concat([ex, syn])
    order = np.argsort(b.pos, kind="stable")
    b = synth._reorder(b, order).validate()
    rt, _ = make_case("cfg1_example", b, L, ref_seq, primers)
    # SURVEY.md section 4 table
    i0 = int(np.flatnonzero(order == 0)[0]); i1 = int(np.flatnonzero(order == 1)[0])
    assert rt[i0]["pos"] == 26 and cigar_string(rt[i0]["cigar"]) == "24S51M76H" and (rt[i0]["ts"], rt[i0]["te"]) == (True, False)
    assert rt[i1]["pos"] == 28254 and cigar_string(rt[i1]["cigar"]) == "31S105M15S" and (rt[i1]["ts"], rt[i1]["te"]) == (False, True)


def cli_case():
    """Run the reference's own run_amplipy (AmpliPy.py:774-963) end to end on SAM files: `aio`, and the
    three-step trim -> variants -> consensus pipeline (config 1's flow).  Outputs are committed as text."""
    import io
    import shutil
    import contextlib
    ref = ref_loader.load_reference()
    d = os.path.join(HERE, "cli")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    L = 4000
    g = synth.random_genome(L, 31)
    primers, amps = synth.make_scheme(L, 12, seed=6, n_alt=2)
    synth.write_bed(os.path.join(d, "primers.bed"), primers, "synth_ref")
    synth.write_fasta(os.path.join(d, "ref.fas"), "synth_ref some description", g)
    b1 = synth.illumina_batch(g, amps, 1200, seed=7, snvs=[(900, "T", 0.6), (2500, "G", 0.04)], p_ins=0.05, p_del=0.05)
    b2 = synth.ont_batch(g, amps, 150, seed=8)
    b = ReadBatch.concat([b1, b2])
    b = synth._reorder(b, np.argsort(b.pos, kind="stable")).validate()
    hdr = "@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:synth_ref\tLN:%d\n@PG\tID:synth\tPN:synth\tVN:1\n" % L
    with open(os.path.join(d, "in.sam"), "w") as f:
        f.write(hdr)
        f.write("\n".join(b.sam_lines(rname="synth_ref")) + "\n")
    argv_saved = list(sys.argv)

    def run(argv, **kw):
        sys.argv[:] = argv
        ref.argv[:] = argv
        with contextlib.redirect_stderr(io.StringIO()):
            ref.run_amplipy(**kw)
    j = lambda n: os.path.join(d, n)
    run(["AmpliPy.py", "aio"], untrimmed_reads_fn=j("in.sam"), primer_fn=j("primers.bed"), reference_fn=j("ref.fas"),
        trimmed_reads_fn=j("aio_trimmed.sam"), variants_fn=j("aio_variants.vcf"), consensus_fn=j("aio_consensus.fas"),
        primer_pos_offset=0, min_length=30, min_quality=20, sliding_window_width=4, min_freq_consensus=0,
        min_freq_variants=0.03, min_depth_consensus=10, min_depth_variants=1, unknown_symbol="N", include_no_primer=False,
        run_trim=True, run_variants=True, run_consensus=True)
    run(["AmpliPy.py", "trim"], untrimmed_reads_fn=j("in.sam"), primer_fn=j("primers.bed"), reference_fn=j("ref.fas"),
        trimmed_reads_fn=j("step_trimmed.sam"), primer_pos_offset=1, min_length=40, min_quality=15, sliding_window_width=5,
        include_no_primer=True, run_trim=True)
    run(["AmpliPy.py", "variants"], trimmed_reads_fn=j("step_trimmed.sam"), reference_fn=j("ref.fas"),
        variants_fn=j("step_variants.vcf"), min_quality=15, min_freq_variants=0.1, min_depth_variants=5, run_variants=True)
    run(["AmpliPy.py", "consensus"], trimmed_reads_fn=j("step_trimmed.sam"), reference_fn=j("ref.fas"),
        consensus_fn=j("step_consensus.fas"), min_quality=15, min_freq_consensus=0.6, min_depth_consensus=4,
        unknown_symbol="x", run_consensus=True)
    sys.argv[:] = argv_saved
    print("cli golden: %d reads -> %s" % (b.n, sorted(os.listdir(d))))


def main():
    cli_case()
    quirk_case()
    example_case()
    L = 6000
    g = synth.random_genome(L, 21)
    # cfg2-like: ARTIC-v3-like Illumina
    primers, amps = synth.make_scheme(L, 19, seed=2)
    b = synth.illumina_batch(g, amps, 5000, seed=2, snvs=[(700, "T", 1.0), (2500, "A", 0.5), (4100, "C", 0.05)])
    make_case("cfg2_illumina", b, L, g, [(s, e) for s, e, _ in primers])
    # cfg3-like: v4.1-like with alt primers, deep coverage on few amplicons, offset 2
    primers, amps = synth.make_scheme(L, 19, seed=3, n_alt=6)
    w = np.zeros(19); w[[3, 4, 5]] = 1 / 3
    b = synth.illumina_batch(g, amps, 5000, seed=3, amp_weights=w, snvs=[(1500, "G", 0.4)])
    make_case("cfg3_deep_alt", b, L, g, [(s, e) for s, e, _ in primers], offset=2, include_no_primer=True)
    # cfg4-like: ONT, lower min_quality so that insertion alleles are exercised
    primers, amps = synth.make_scheme(L, 19, seed=4)
    b = synth.ont_batch(g, amps, 1200, seed=4)
    make_case("cfg4_ont", b, L, g, [(s, e) for s, e, _ in primers], min_freq_variants=0.1)
    make_case("cfg4_ont_mq10", b, L, g, [(s, e) for s, e, _ in primers], min_quality=10, window=6)
    # adversarial CIGAR/quality fuzz incl. offsets, window widths, min_quality 0/30
    Lf = 3000
    gf = synth.random_genome(Lf, 5)
    primers, amps = synth.make_scheme(Lf, 9, amp_len=350, seed=3, n_alt=2)
    for seed in range(4):
        recs = synth.fuzz_records(Lf, 1200, seed=seed, ont_like=(seed % 2 == 1))
        b = ReadBatch.from_records(recs)
        make_case("fuzz%d" % seed, b, Lf, gf, [(s, e) for s, e, _ in primers], offset=[0, 2, 5, 10][seed],
                  min_quality=[20, 0, 30, 20][seed], window=[4, 1, 10, 4][seed], include_no_primer=(seed % 2 == 0),
                  min_depth_consensus=[10, 1, 3, 10][seed], min_freq_consensus=[0.0, 0.5, 0.0, 0.9][seed],
                  min_depth_variants=[1, 5, 1, 20][seed], min_freq_variants=[0.03, 0.2, 0.0, 0.03][seed],
                  unknown_symbol="Nn?N"[seed])


if __name__ == "__main__":
    main()
