"""GPU (B200): parity of the CUDA path, called through the C ABI, against (1) the golden fixtures
produced by the unmodified reference and (2) the oracle on larger seeded inputs; plus
size-independent properties at the sizes of configs 2-5 (1 M reads; 20 M-read depth; 500 k ONT-like
reads; a 384-sample plate)."""
import numpy as np
import pytest

import golden_io
import parity
from amplipy_b200 import calling, synth
from amplipy_b200.batch import ReadBatch
from amplipy_b200.primers import find_overlapping_primers, max_primer_len

pytestmark = pytest.mark.gpu
CASES = golden_io.list_cases()


def make_engine(**kw):
    from amplipy_b200.engine import Engine
    return Engine(**kw)


class DeviceEngine:
    """Engine driven through the device-resident entry point (amp_process_device)."""

    def __init__(self, **kw):
        self.e = make_engine(**kw)

    def process(self, batch, trim=True, pileup=True, sample=0):
        import torch
        d = self.e.upload(batch, trim_out=trim)
        self.e.process_device(d, trim=trim, pileup=pileup, sample=sample)
        torch.cuda.synchronize()
        return self.e.download_trim(batch, d) if trim else None

    def __getattr__(self, k):
        return getattr(self.e, k)


@pytest.mark.parametrize("name", CASES)
def test_golden_aio_host_path(name):
    parity.check_case_aio(make_engine, name)


@pytest.mark.parametrize("name", CASES)
def test_golden_aio_device_path(name):
    parity.check_case_aio(DeviceEngine, name)


@pytest.mark.parametrize("name", CASES)
def test_golden_pileup_only(name):
    parity.check_case_pileup_only(make_engine, name)


@pytest.mark.parametrize("name", ["cfg1_example", "cfg2_illumina", "fuzz1"])
def test_golden_trim_then_variants_pipeline(name):
    parity.check_case_pipeline(make_engine, name)


def _scheme(L=29903, n_amp=98, seed=2, n_alt=0):
    g = synth.random_genome(L, 7)
    primers, amps = synth.make_scheme(L, n_amp, seed=seed, n_alt=n_alt)
    prim = [(s, e) for s, e, _ in primers]
    return g, prim, amps


def _against_oracle(oracle, b, g, prim, **kw):
    return parity.check_against_oracle(make_engine, oracle, b, g, prim, **kw)


def test_illumina_300k_vs_oracle(oracle_lib):
    """> one host chunk (262144 reads), so the chunked H2D/D2H pipeline and pointer rebasing are covered."""
    g, prim, amps = _scheme()
    b = synth.illumina_batch(g, amps, 300_000, seed=21, snvs=[(1000, "T", 0.5), (20000, "A", 0.03)])
    _against_oracle(oracle_lib, b, g, prim)


def test_illumina_alt_primers_offset_vs_oracle(oracle_lib):
    g, prim, amps = _scheme(seed=3, n_alt=20)
    w = np.zeros(98); w[[10, 11, 50]] = 1 / 3          # deep coverage on three amplicons (atomic contention)
    b = synth.illumina_batch(g, amps, 120_000, seed=22, amp_weights=w)
    _against_oracle(oracle_lib, b, g, prim, offset=3, inc=True)


def test_ont_40k_vs_oracle(oracle_lib):
    g, prim, amps = _scheme(seed=4)
    b = synth.ont_batch(g, amps, 40_000, seed=23)
    _against_oracle(oracle_lib, b, g, prim, ins_slots=1 << 22, arena=256 << 20)


def test_ont_low_min_quality_vs_oracle(oracle_lib):
    g, prim, amps = _scheme(seed=4)
    b = synth.ont_batch(g, amps, 15_000, seed=24)
    _against_oracle(oracle_lib, b, g, prim, mq=7, w=6, ins_slots=1 << 22, arena=256 << 20)


def test_unsorted_input_vs_oracle(oracle_lib):
    """Order-agnostic like the reference: shuffled reads force a window flush on almost every tile."""
    g, prim, amps = _scheme()
    b = synth.illumina_batch(g, amps, 60_000, seed=25, sort=False)
    _against_oracle(oracle_lib, b, g, prim)


def test_fuzz_vs_oracle(oracle_lib):
    L = 4000
    g = synth.random_genome(L, 5)
    primers, _ = synth.make_scheme(L, 11, amp_len=350, seed=3, n_alt=3)
    prim = [(s, e) for s, e, _ in primers]
    for seed in range(40, 46):
        b = ReadBatch.from_records(synth.fuzz_records(L, 3000, seed=seed, ont_like=bool(seed & 1)))
        _against_oracle(oracle_lib, b, g, prim, offset=[0, 2, 7][seed % 3], mq=[20, 0, 33][seed % 3], w=[4, 1, 9][seed % 3],
                        inc=bool(seed & 2))


def test_reads_at_the_ends_of_the_reference_vs_oracle(oracle_lib):
    """Reads at position 0 / ending on the last base, a primer that reaches the last base: a read swallowed by the start clip
    ends up at pos == L with an all-S CIGAR; the reference piles it up without touching anything and so must the kernels."""
    L = 240
    g = synth.random_genome(L, 3)
    prim = [(0, 22), (60, 84), (150, 171), (L - 24, L)]
    for seed in (111, 112):
        recs = synth.fuzz_records(L, 3000, seed=seed, max_len=40, edges=0.5)
        recs += [(L - 2, 0, 0, [(0, 2)], "AC", [30, 30]), (L - 6, 16, 0, [(0, 6)], "ACGTAC", [30] * 6)]
        b = ReadBatch.from_records(sorted(recs, key=lambda r: r[0]))
        _against_oracle(oracle_lib, b, g, prim, offset=seed % 3, ml=5)


def test_single_indel_reads_vs_oracle(oracle_lib):
    """[H][S] M (I|D) M [S][H] reads finished in registers next to [S]M[S]: heavy indel / clip rates and quality patterns that
    move the quality clip across the indel on both strands."""
    g, prim, amps = _scheme(seed=3, n_alt=5)
    rng = np.random.default_rng(9)
    b = synth.illumina_batch(g, amps, 200_000, seed=41, p_ins=0.3, p_del=0.3, p_clip=0.3, p_hard=0.1, p_short=0.2)
    q = b.qual
    for i in range(0, b.n, 3):                       # weak windows / sprinkled weak bases on every third read
        lo, hi = int(b.qual_off[i]), int(b.qual_off[i + 1])
        if hi - lo < 12:
            continue
        if i % 2:
            p0 = int(rng.integers(lo, hi - 4)); q[p0:p0 + 4] = rng.integers(0, 25, 4)
        else:
            k = int(rng.integers(1, 12)); q[rng.integers(lo, hi, k)] = rng.integers(0, 20, k)
    _against_oracle(oracle_lib, b, g, prim)
    _against_oracle(oracle_lib, b, g, prim, mq=30, offset=2)


def test_empty_and_tiny_batches():
    g, prim, amps = _scheme(L=3000, n_amp=9)
    tables = find_overlapping_primers(3000, prim, 0)
    eng = make_engine(ref_len=3000, primer_tables=tables, max_primer_len=max_primer_len(prim))
    empty = ReadBatch.from_records([])
    t = eng.process(empty)
    assert t.pos.shape == (0,) and not eng.counts().any() and eng.insertions().k == 0
    one = ReadBatch.from_records([(100, 0, 0, [(0, 50)], "ACGTN" * 10, [30] * 50)])
    eng.process(one)
    assert int(eng.counts().sum()) == 50
    res = eng.call(g)
    assert int(res.depth.sum()) == 50


def test_full_size_properties_cfg2():
    """Config 2 at full size (1M reads): order independence, additivity over read subsets, host path ==
    device path, and the depth identity sum(counts) + sum(insertion counts) == sum(depth)."""
    import torch
    g, prim, amps = _scheme()
    L = len(g)
    b = synth.illumina_batch(g, amps, 1_000_000, seed=2)
    tables = find_overlapping_primers(L, prim, 0)
    mk = lambda: make_engine(ref_len=L, primer_tables=tables, max_primer_len=max_primer_len(prim))
    e1 = mk()
    t1 = e1.process(b)
    c1 = e1.counts(); i1 = e1.insertions().as_dict()
    # device-resident path
    e2 = mk()
    d = e2.upload(b)
    e2.process_device(d)
    torch.cuda.synchronize()
    t2 = e2.download_trim(b, d)
    assert np.array_equal(e2.counts(), c1) and e2.insertions().as_dict() == i1
    assert np.array_equal(t1.pos, t2.pos) and np.array_equal(t1.flags, t2.flags) and np.array_equal(t1.ncig, t2.ncig)
    # additivity: two halves accumulated into one context, processed in reverse order
    e3 = mk()
    h = b.n // 2
    e3.process(b, first=h, n=b.n - h)
    e3.process(b, first=0, n=h)
    assert np.array_equal(e3.counts(), c1) and e3.insertions().as_dict() == i1
    res = e1.call(g)
    assert int(res.depth.astype(np.int64).sum()) == int(c1.astype(np.int64).sum()) + sum(i1.values())
    # every kept read satisfies the write gate on its own outputs
    assert ((t1.flags & 8) != 0).sum() > 0
    assert e1.error_flags() == 0


def test_multi_sample_context_and_insertion_merge(oracle_lib):
    """Plate mode: independent count matrices in one context; and merging another context's insertion
    table (the multi-GPU exchange step) reproduces the single-context result."""
    g, prim, amps = _scheme(L=6000, n_amp=19)
    L = len(g)
    tables = find_overlapping_primers(L, prim, 0)
    bs = [synth.illumina_batch(g, amps, 20_000, seed=100 + i, p_ins=0.2) for i in range(3)]
    plate = make_engine(ref_len=L, primer_tables=tables, max_primer_len=max_primer_len(prim), n_samples=3)
    for i, b in enumerate(bs):
        plate.process(b, sample=i)
    pins = plate.insertions()
    for i, b in enumerate(bs):
        solo = make_engine(ref_len=L, primer_tables=tables, max_primer_len=max_primer_len(prim))
        solo.process(b)
        assert np.array_equal(solo.counts(), plate.counts(i))
        assert solo.insertions().as_dict() == pins.as_dict(i)
    # merge: ctx A sees half the reads, ctx B the other half; B's table merged into A equals the whole
    b = bs[0]
    a_ = make_engine(ref_len=L, primer_tables=tables, max_primer_len=max_primer_len(prim))
    b_ = make_engine(ref_len=L, primer_tables=tables, max_primer_len=max_primer_len(prim))
    a_.process(b, first=0, n=b.n // 2)
    b_.process(b, first=b.n // 2, n=b.n - b.n // 2)
    ib = b_.insertions()
    a_.merge_insertions(ib.sample, ib.pos, ib.count, ib.str_off, ib.chars)
    assert a_.insertions().as_dict() == pins.as_dict(0)
    assert a_.error_flags() == 0


def test_multi_gpu_sharding_nccl():
    """Needs >= 2 visible GPUs (gpurun --gpus 2); skipped on single-GPU boxes."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("single GPU")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)),
                        "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(root, "tests", "dist_gpu_check.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "dist_gpu_check ok" in r.stdout, r.stdout[-3000:]


def _long_read_batch(L, n, seed, read_len=(3000, 9000)):
    """PacBio/ONT-like long reads: rows far larger than the staging buffers, hundreds of CIGAR ops (global scratch
    rows), deletions/insertions/skips everywhere."""
    rng = np.random.default_rng(seed)
    g = synth.random_genome(L, seed)
    recs = []
    for _ in range(n):
        target = int(rng.integers(*read_len))
        ops = []
        if rng.random() < 0.5:
            ops.append((4, int(rng.integers(1, 60))))
        rl = 0
        while rl < target:
            m = int(rng.integers(5, 120))
            ops.append((0, m)); rl += m
            r = rng.random()
            if r < 0.4:
                ops.append((1, int(rng.integers(1, 6))))
            elif r < 0.8:
                d = int(rng.integers(1, 8)); ops.append((2, d)); rl += d
        if ops[-1][0] != 0:
            ops.append((0, 10)); rl += 10
        if rng.random() < 0.5:
            ops.append((4, int(rng.integers(1, 60))))
        pos = int(rng.integers(1, L - rl - 2))
        ql = sum(x for op, x in ops if op in (0, 1, 4))
        seq = "".join(rng.choice(list("ACGT"), size=ql))
        qual = rng.integers(5, 40, ql).tolist()
        recs.append((pos, int(rng.choice([0, 16])), 0, ops, seq, qual))
    recs.sort(key=lambda r: r[0])
    return g, ReadBatch.from_records(recs)


def test_long_reads_vs_oracle(oracle_lib):
    L = 30000
    g, b = _long_read_batch(L, 300, seed=77)
    primers, _ = synth.make_scheme(L, 60, amp_len=500, seed=9)
    prim = [(s, e) for s, e, _ in primers]
    assert int(b.n_cigar.max()) > 100 and int(b.l_seq.max()) > 5000
    _against_oracle(oracle_lib, b, g, prim, mq=15, ins_slots=1 << 20, arena=64 << 20)


def test_longer_short_reads_vs_oracle(oracle_lib):
    """250- and 301-bp reads through the warp-autonomous kernel: fewer reads per batch, aligned runs that need all
    thirty-two 8-window blocks, and runs beyond them that take the generic phase."""
    g, prim, amps = _scheme(seed=5)
    for n, seed, rl in ((60_000, 51, 250), (30_000, 52, 301)):
        b = synth.illumina_batch(g, amps, n, seed=seed, read_len=rl)
        _against_oracle(oracle_lib, b, g, prim)


def _ins_table_arrays(ins, sample=None):
    """Canonical array form of an insertion table: rows (sample, pos, len, count, FNV of the text), sorted."""
    off = np.asarray(ins.str_off, np.int64)
    ln = np.diff(off)
    h = np.full(len(ln), 0xCBF29CE484222325, np.uint64)
    ch = np.asarray(ins.chars, np.uint8)
    with np.errstate(over="ignore"):
        for j in range(int(ln.max()) if len(ln) else 0):      # column-wise over the strings (vectorised over alleles)
            m = ln > j
            h[m] = (h[m] ^ ch[off[:-1][m] + j].astype(np.uint64)) * np.uint64(0x100000001B3)
    rows = np.stack([ins.sample.astype(np.int64), ins.pos.astype(np.int64), ln, ins.count.astype(np.int64), h.view(np.int64)], 1)
    if sample is not None:
        rows = rows[rows[:, 0] == sample]
    return rows[np.lexsort(rows.T[::-1])]


def test_full_depth_properties_cfg3():
    """Config 3's shape (v4.1-like scheme with alt primers, 20 M reads, ~100,000x): one million seeded reads accumulated twenty
    times into one context.  Counts and insertion counts scale by exactly 20 (no saturation / lost updates at that
    depth), and calling is scale-invariant: count/total is the same rational, so the float64 frequencies, the allele
    order and the consensus are bit-identical to the 1x sample."""
    g, prim, amps = _scheme(seed=3, n_alt=10)
    L = len(g)
    b = synth.illumina_batch(g, amps, 1_000_000, seed=3)
    tables = find_overlapping_primers(L, prim, 0)
    mk = lambda: make_engine(ref_len=L, primer_tables=tables, max_primer_len=max_primer_len(prim))
    e1 = mk()
    e1.process(b)
    c1 = e1.counts().astype(np.int64); i1 = _ins_table_arrays(e1.insertions())
    r1 = e1.call(g, min_depth_consensus=10, min_depth_variants=1)
    e20 = mk()
    d = e20.upload(b)
    for _ in range(20):
        e20.process_device(d)
    import torch
    torch.cuda.synchronize()
    c20 = e20.counts().astype(np.int64); i20 = _ins_table_arrays(e20.insertions())
    assert int(c20.max()) > 50_000                                   # the depth the config is about
    assert np.array_equal(c20, 20 * c1)
    assert np.array_equal(i20[:, [0, 1, 2, 4]], i1[:, [0, 1, 2, 4]]) and np.array_equal(i20[:, 3], 20 * i1[:, 3])
    r20 = e20.call(g, min_depth_consensus=10, min_depth_variants=1)
    assert np.array_equal(r20.depth.astype(np.int64), 20 * r1.depth.astype(np.int64))
    assert np.array_equal(r20.fixed_freq, r1.fixed_freq) and np.array_equal(r20.fixed_rank, r1.fixed_rank)
    deep = r1.top_count >= 10                                        # consensus gate passes in both
    assert np.array_equal(r20.top_id[deep], r1.top_id[deep]) and np.array_equal(r20.alt_mask, r1.alt_mask)
    assert e20.error_flags() == 0


def test_ont_properties_cfg4():
    """Config 4's shape (ONT-like, indel-rich) at 500 k reads: host path == device path, accumulation in halves in reverse
    order == one pass (counts and the whole insertion table), and the depth identity."""
    import torch
    g, prim, amps = _scheme()
    L = len(g)
    b = synth.ont_batch(g, amps, 500_000, seed=4)
    tables = find_overlapping_primers(L, prim, 0)
    mk = lambda: make_engine(ref_len=L, primer_tables=tables, max_primer_len=max_primer_len(prim), ins_slots=1 << 25,
                             ins_arena_bytes=2 << 30)
    e1 = mk()
    t1 = e1.process(b)
    c1 = e1.counts(); i1 = _ins_table_arrays(e1.insertions())
    assert len(i1) > 100_000 and e1.error_flags() == 0
    res = e1.call(g)
    assert int(res.depth.astype(np.int64).sum()) == int(c1.astype(np.int64).sum()) + int(i1[:, 3].sum())
    del e1
    e2 = mk()
    d = e2.upload(b)
    h = b.n // 2
    e2.process_device(d, first=h, n=b.n - h)
    e2.process_device(d, first=0, n=h)
    torch.cuda.synchronize()
    t2 = e2.download_trim(b, d)
    assert np.array_equal(e2.counts(), c1) and np.array_equal(_ins_table_arrays(e2.insertions()), i1)
    assert np.array_equal(t1.pos, t2.pos) and np.array_equal(t1.flags, t2.flags) and np.array_equal(t1.ncig, t2.ncig)
    for i in range(0, b.n, 257):
        assert t1.cigartuples(i) == t2.cigartuples(i)
    assert e2.error_flags() == 0


def test_plate_384_samples_cfg5():
    """Config 5's shape: 384 samples in one context (reduced to 3,000 reads per sample).  Every sample's count matrix and
    insertion alleles equal a solo run of that sample (spot-checked), and one calling launch over the plate equals the
    solo calls."""
    g, prim, amps = _scheme()
    L = len(g)
    tables = find_overlapping_primers(L, prim, 0)
    S = 384
    plate = make_engine(ref_len=L, primer_tables=tables, max_primer_len=max_primer_len(prim), n_samples=S)
    batches = {}
    for i in range(S):
        bi = synth.illumina_batch(g, amps, 3_000, seed=1000 + i, snvs=[(500 + 70 * i, "T", 0.6)])
        plate.process(bi, sample=i)
        if i in (0, 1, 191, 383):
            batches[i] = bi
    pins = plate.insertions()
    pres = plate.call(g)
    assert pres.depth.shape == (S * L,)
    for i, bi in batches.items():
        solo = make_engine(ref_len=L, primer_tables=tables, max_primer_len=max_primer_len(prim))
        solo.process(bi)
        assert np.array_equal(solo.counts(), plate.counts(i))
        assert solo.insertions().as_dict() == pins.as_dict(i)
        sres = solo.call(g)
        sl = slice(i * L, (i + 1) * L)
        for f in ("depth", "top_count", "pos_flags", "ref_count", "fixed_freq", "fixed_rank", "alt_mask"):
            assert np.array_equal(getattr(pres, f)[sl], getattr(sres, f)), f
        fixed = sres.top_id < 6                                       # insertion ids are table-order dependent
        assert np.array_equal(pres.top_id[sl][fixed], sres.top_id[fixed])
    assert plate.error_flags() == 0


# ---- BAM decoded on the device (amp_bam_decode_host / amp_process_decoded) -------------------------------------------------------
def _bam_bytes(tmp_path, b, L, level=6):
    import os
    from amplipy_b200 import alnio
    path = os.path.join(str(tmp_path), "in_%d.bam" % level)
    alnio.write_bam(path, "@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:ref\tLN:%d\n@PG\tID:synth\tPN:synth\n" % L, [("ref", L)], b, level=level)
    return open(path, "rb").read()


@pytest.mark.parametrize("level", [1, 6, 9])
def test_device_bam_decode_equals_host_decoder(tmp_path, level):
    """Byte-equal struct-of-arrays batch (and record offsets) from the device-side inflate + record scatter and from the host
    decoder (zlib + amp_bam_fill), on Illumina-like and ONT-like records, at several deflate levels."""
    from amplipy_b200 import alnio
    g, prim, amps = _scheme(L=6000, n_amp=18)
    for b in (synth.illumina_batch(g, amps, 30_000, seed=51, p_ins=0.1, p_del=0.1), synth.ont_batch(g, amps, 2_000, seed=52)):
        raw = _bam_bytes(tmp_path, b, 6000, level)
        a = alnio._read_bam(raw)
        eng = make_engine(ref_len=6000)
        info = eng.decode_bam(raw, alnio.bam_layout(raw))
        assert info["n"] == b.n and info["sum_cig"] == int(b.cig_off[-1]) and info["sum_qual"] == int(b.qual_off[-1])
        got, rec_off = eng.decoded_batch()
        for f in ("pos", "flag", "tlen", "cig_off", "cigar", "seq_off", "seq", "qual_off", "qual"):
            assert np.array_equal(getattr(got, f), getattr(a.batch, f)), f
            assert np.array_equal(getattr(got, f), getattr(b, f)), f
        assert np.array_equal(rec_off.astype(np.int64), a.bam_rec_off)


def test_device_bam_path_vs_oracle(oracle_lib, tmp_path):
    """File bytes -> device decode -> fused kernel -> calling, against the oracle on the same reads."""
    from amplipy_b200 import alnio
    g, prim, amps = _scheme()
    b = synth.illumina_batch(g, amps, 120_000, seed=53, snvs=[(1000, "T", 0.5)])
    raw = _bam_bytes(tmp_path, b, len(g))

    class BamEngine:
        def __init__(self, **kw):
            self.e = make_engine(**kw)

        def process(self, batch, trim=True, pileup=True, sample=0):
            from amplipy_b200.engine import TrimResult
            self.e.decode_bam(raw, alnio.bam_layout(raw))
            out = self.e.process_decoded(trim=trim, pileup=pileup, sample=sample)
            return TrimResult(batch, *out) if trim else None

        def __getattr__(self, k):
            return getattr(self.e, k)
    parity.check_against_oracle(BamEngine, oracle_lib, b, g, prim)


def test_device_bam_decode_rejects_bad_input(tmp_path):
    from amplipy_b200 import alnio
    from amplipy_b200.engine import AmpError
    g, prim, amps = _scheme(L=6000, n_amp=18)
    b = synth.illumina_batch(g, amps, 5_000, seed=54)
    raw = _bam_bytes(tmp_path, b, 6000)
    lay = alnio.bam_layout(raw)
    eng = make_engine(ref_len=6000)
    bad = bytearray(raw)
    k = len(lay["in_off"]) // 2
    a = int(lay["in_off"][k])
    for i in range(a + 40, a + 60):
        bad[i] ^= 0x5A                                    # garble one block's deflate stream
    with pytest.raises(AmpError):
        eng.decode_bam(bytes(bad), lay)
    # records straddling block boundaries (plain 0xff00-byte blocks, as a non-htslib writer may produce them)
    payload = alnio.bgzf_decompress(raw).tobytes()
    raw2 = alnio.bgzf_compress(payload)
    with pytest.raises(AmpError):
        eng.decode_bam(raw2, alnio.bam_layout(raw2))
    info = eng.decode_bam(raw, lay)                       # the context is still usable
    assert info["n"] == b.n


# ---- function-level drop-ins with the reference's signatures (amplipy_b200/compat.py) -------------------------------------------
class _Seg:
    """The pysam.AlignedSegment attributes the reference's functions touch."""

    def __init__(self, rec):
        pos, flag, tlen, ops, seq, qual = rec
        self.reference_start, self.flag, self.template_length = pos, flag, tlen
        self.cigartuples, self.query_sequence, self.query_qualities = [tuple(x) for x in ops], seq, list(qual)


def test_function_level_dropins_match_the_golden_quirks():
    """trim_read(s, ...), update_base_counts(counts, s, mq), alleles_from_counts(dict) with the reference's signatures, read by
    read, against the quirk vectors produced by the unmodified reference (tests/golden/quirks.npz)."""
    from amplipy_b200 import compat
    b, meta, arr = golden_io.load_case("quirks")
    p = meta["params"]
    L = meta["L"]
    prim = [tuple(x) for x in meta["primers"]]
    mn, mx = compat.find_overlapping_primers(L, prim, p["offset"])
    counts = [{'A': 0, 'C': 0, 'G': 0, 'T': 0, 'N': 0, '-': 0} for _ in range(L)]
    mpl = max(e - s for s, e in prim)
    for i in range(b.n):
        s = _Seg(b.record(i))
        if (s.flag & 4) or not s.cigartuples:
            continue
        ts, te, tq = compat.trim_read(s, mn, mx, mpl, p["min_quality"], p["window"])
        fl = int(arr["t_flags"][i])
        assert (ts, te, tq) == (bool(fl & 1), bool(fl & 2), bool(fl & 4)), i
        assert s.reference_start == int(arr["t_pos"][i])
        a = int(b.cig_off[i]) + 3 * i
        assert [((n << 4) | op) for op, n in s.cigartuples] == arr["t_cigar"][a:a + int(arr["t_ncig"][i])].tolist(), i
        compat.update_base_counts(counts, s, p["min_quality"])
    want = arr["counts_aio"]
    for ch in range(6):
        assert [counts[q]["ACGTN-"[ch]] for q in range(L)] == want[ch].tolist()
    got_ins = {(q, k): v for q in range(L) for k, v in counts[q].items() if not (len(k) == 1 and k in "ACGTN-")}
    assert got_ins == golden_io.ins_from_meta(meta)
    al_off = arr["al_off"]
    for q in range(L):
        total, al = compat.alleles_from_counts(counts[q])
        a0, a1 = int(al_off[q]), int(al_off[q + 1])
        assert total == int(arr["depth_aio"][q])
        assert [x[2] for x in al] == meta["al_sym"][a0:a1] and [x[0] for x in al] == arr["al_count"][a0:a1].tolist()
        assert [x[1] for x in al] == arr["al_freq"][a0:a1].tolist()


# ---- heterogeneous plates: a primer scheme / reference per sample in one context ---------------------------------
def test_device_primer_tables_equal_find_overlapping_primers():
    rng = np.random.default_rng(17)
    L = 5000
    eng = make_engine(ref_len=L, n_samples=4)
    for smp, (n, off) in enumerate([(40, 0), (200, 3), (1, 10), (0, 0)]):
        st = np.sort(rng.integers(0, L - 40, n))
        prim = [(int(s), int(s + rng.integers(15, 35))) for s in st]
        eng.set_scheme(smp, prim, off)
        mn, mx, mpl = eng.get_scheme(smp)
        wmn, wmx = find_overlapping_primers(L, prim, off) if n else (np.full(L, -1, np.int32), np.full(L, -1, np.int32))
        assert np.array_equal(mn, wmn) and np.array_equal(mx, wmx)
        assert mpl == (max_primer_len(prim) if n else 0)


def test_mixed_plate_in_one_context_equals_solo_runs():
    """Three samples with different primer schemes, offsets, references and reference lengths in ONE context: per-sample trim
    outputs, counts, insertions and calling results equal those of a context of their own; one calling launch covers the plate."""
    specs = [(29903, 98, 2, 0, 61), (12000, 40, 3, 2, 62), (29903, 98, 4, 0, 63)]
    Lmax = max(s[0] for s in specs)
    plate = make_engine(ref_len=Lmax, n_samples=len(specs))
    solo_res = []
    batches = []
    for smp, (L, n_amp, seed, off, bseed) in enumerate(specs):
        g = synth.random_genome(L, 10 + smp)
        primers, amps = synth.make_scheme(L, n_amp, seed=seed)
        prim = [(s, e) for s, e, _ in primers]
        b = synth.illumina_batch(g, amps, 30_000, seed=bseed, p_ins=0.05, snvs=[(L // 3, "T", 0.4)])
        batches.append((b, g, prim, off, L))
        solo = make_engine(ref_len=L, primer_tables=find_overlapping_primers(L, prim, off), max_primer_len=max_primer_len(prim))
        t = solo.process(b)
        ins = solo.insertions()
        res = solo.call(g)
        solo_res.append((t, solo.counts(), ins.as_dict(), calling.consensus_string(res, ins), calling.variant_records(res, ins, g, solo.counts())))
        plate.set_scheme(smp, prim, off, ref_len=L)
        plate.set_sample_reference(smp, g)
    outs = [plate.process(b, sample=smp) for smp, (b, g, prim, off, L) in enumerate(batches)]
    assert plate.error_flags() == 0
    ins = plate.insertions()
    res = plate.call(None)
    for smp, (b, g, prim, off, L) in enumerate(batches):
        t, counts, insd, cons, recs = solo_res[smp]
        assert np.array_equal(outs[smp].pos, t.pos) and np.array_equal(outs[smp].flags, t.flags) and np.array_equal(outs[smp].ncig, t.ncig)
        assert np.array_equal(plate.counts(smp)[:, :L], counts)
        mine = {(int(p), s): int(c) for sm, p, c, s in zip(ins.sample, ins.pos, ins.count, ins.strs) if sm == smp}
        assert mine == insd
        assert calling.consensus_string(res, ins, smp)[:L] == cons
        got = calling.variant_records(res, ins, g + "N" * (Lmax - L), plate.counts(smp), smp)
        assert len(got) == len(recs)
        for a, c in zip(got, recs):
            assert tuple(a[:6]) == tuple(c[:6]) and a[6] == c[6] and a[7] == c[7] and tuple(a[8]) == tuple(c[8])


# ---- BGZF deflate on the device (amp_bgzf_deflate_host) ----------------------------------------------------------------------------
def _bgzf_members(buf):
    """(payload bytes, isize) of every BGZF block of buf, checking the fixed header fields"""
    out, o = [], 0
    while o < len(buf):
        assert buf[o:o + 4] == b"\x1f\x8b\x08\x04" and buf[o + 12:o + 16] == b"BC\x02\x00"
        bsize = int.from_bytes(buf[o + 16:o + 18], "little") + 1
        out.append((buf[o + 18:o + bsize - 8], int.from_bytes(buf[o + bsize - 4:o + bsize], "little")))
        o += bsize
    assert o == len(buf)
    return out


@pytest.mark.gpu
def test_device_bgzf_deflate_round_trip(tmp_path):
    """Whatever the device compressor writes is a BGZF file: every block a gzip member with the right CRC-32 / ISIZE (Python's gzip
    checks both), the EOF block at the end, blocks cut where the caller said; incompressible blocks come back stored."""
    import gzip
    from amplipy_b200 import alnio
    g, prim, amps = _scheme(L=6000, n_amp=18)
    b = synth.illumina_batch(g, amps, 40_000, seed=61, p_ins=0.05, p_del=0.05)
    raw = _bam_bytes(tmp_path, b, 6000)
    data = alnio.bgzf_decompress(raw)
    rng = np.random.default_rng(5)
    noise = rng.integers(0, 256, 200_000, dtype=np.uint8)
    eng = make_engine(ref_len=6000)
    for payload, step in ((data, 0xff00), (data[:1_000_003], 40_000), (noise, 0xff00), (np.concatenate([data[:100_000], noise, data[:70_001]]), 65_000),
                          (data[:10], 0xff00), (data[:0], 0xff00)):
        bstart = np.unique(np.concatenate([np.arange(0, payload.size, step), [payload.size]])).astype(np.int64)
        if payload.size == 0:
            bstart = np.zeros(1, np.int64)
        out = eng.bgzf_deflate(payload, bstart).tobytes()
        assert out.endswith(bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0]))
        assert gzip.decompress(out) == payload.tobytes()
        members = _bgzf_members(out)
        assert [m[1] for m in members[:-1]] == list(np.diff(bstart)) and members[-1][1] == 0
        assert alnio.bgzf_decompress(out).tobytes() == payload.tobytes()           # the repo's own host inflater
    # compressible data does get smaller, noise does not grow by more than the block overhead
    bs = np.unique(np.concatenate([np.arange(0, data.size, 0xff00), [data.size]])).astype(np.int64)
    assert eng.bgzf_deflate(data, bs).size < 0.45 * data.size
    bn = np.unique(np.concatenate([np.arange(0, noise.size, 0xff00), [noise.size]])).astype(np.int64)
    assert eng.bgzf_deflate(noise, bn).size <= noise.size + 31 * (bn.size - 1) + 28
    # and the device decoder reads it back: file bytes -> amp_bam_decode_host -> the same batch
    cuts = np.unique(np.concatenate([[0], np.cumsum(alnio.bam_layout(raw)["out_len"].astype(np.int64))]))     # the record-aligned blocks of the input
    out = eng.bgzf_deflate(data, cuts).tobytes()
    info = eng.decode_bam(out, alnio.bam_layout(out))
    got, _ = eng.decoded_batch()
    assert info["n"] == b.n
    for f in ("pos", "flag", "tlen", "cig_off", "cigar", "seq_off", "seq", "qual_off", "qual"):
        assert np.array_equal(getattr(got, f), getattr(b, f)), f


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["illumina", "ont", "nothing_kept"])
def test_cli_bam_output_device_vs_zlib(tmp_path, monkeypatch, kind):
    """`trim` BAM -> BAM with the records rebuilt and compressed on the device (default: amp_decoded_write_bam) and on the host with
    zlib level 6 (AMPLIPY_BAM_LEVEL): the same records, byte for byte."""
    import os
    from amplipy_b200 import alnio, cli
    g, prim, amps = _scheme(L=6000, n_amp=18)
    if kind == "ont":
        b = synth.ont_batch(g, amps, 6_000, seed=63)
    else:
        b = synth.illumina_batch(g, amps, 60_000 if kind == "illumina" else 3_000, seed=62, p_ins=0.05, p_del=0.05)
    d = str(tmp_path)
    open(os.path.join(d, "in.bam"), "wb").write(_bam_bytes(tmp_path, b, 6000))
    with open(os.path.join(d, "p.bed"), "w") as f:
        for k, (s, e) in enumerate(prim):
            f.write("ref\t%d\t%d\tp%d\n" % (s, e, k))
    with open(os.path.join(d, "ref.fas"), "w") as f:
        f.write(">ref\n%s\n" % g)
    j = lambda n: os.path.join(d, n)
    monkeypatch.delenv("AMPLIPY_BAM_LEVEL", raising=False)
    extra = ["-ml", "100000"] if kind == "nothing_kept" else []                      # a minimum length no read reaches: nothing passes the gate
    cli.main(["trim", "-i", j("in.bam"), "-p", j("p.bed"), "-r", j("ref.fas"), "-o", j("dev.bam")] + extra)
    monkeypatch.setenv("AMPLIPY_BAM_LEVEL", "6")
    cli.main(["trim", "-i", j("in.bam"), "-p", j("p.bed"), "-r", j("ref.fas"), "-o", j("zlib.bam")] + extra)
    a, z = alnio.read_alignments(j("dev.bam")), alnio.read_alignments(j("zlib.bam"))
    assert a.n == z.n and (a.n > 0) == (kind != "nothing_kept") and a.header_text.splitlines()[:-1] == z.header_text.splitlines()[:-1]   # (@PG quotes the command)
    for f in ("pos", "flag", "tlen", "cig_off", "cigar", "seq_off", "seq", "qual_off", "qual"):
        assert np.array_equal(getattr(a.batch, f), getattr(z.batch, f)), f
    import gzip
    ra, rz = gzip.decompress(open(j("dev.bam"), "rb").read()), gzip.decompress(open(j("zlib.bam"), "rb").read())
    assert ra[alnio.bam_layout(open(j("dev.bam"), "rb").read())["body_off"]:] == rz[alnio.bam_layout(open(j("zlib.bam"), "rb").read())["body_off"]:]
    assert os.path.getsize(j("dev.bam")) < 2 * os.path.getsize(j("zlib.bam")) + 200
    eng_cls = __import__("amplipy_b200.engine", fromlist=["Engine"]).Engine
    assert hasattr(eng_cls, "decoded_write_bam")


@pytest.mark.gpu
def test_device_writer_error_paths(tmp_path):
    """Argument and state errors of the two output entry points come back as AmpError, and the context stays usable."""
    from amplipy_b200 import alnio
    from amplipy_b200.engine import AmpError
    g, prim, amps = _scheme(L=6000, n_amp=18)
    b = synth.illumina_batch(g, amps, 4_000, seed=71)
    raw = _bam_bytes(tmp_path, b, 6000)
    lay = alnio.bam_layout(raw)
    head = alnio._bam_header_bytes(lay["header_text"], lay["refs"])
    from amplipy_b200.primers import find_overlapping_primers, max_primer_len
    eng = make_engine(ref_len=6000, primer_tables=find_overlapping_primers(6000, prim, 0), max_primer_len=max_primer_len(prim))
    with pytest.raises(AmpError):
        eng._decoded = {"n": 0, "sum_cig": 0, "sum_seq": 0, "sum_qual": 0, "raw_bytes": 0}
        eng.decoded_write_bam(head)                                   # nothing decoded yet
    eng.decode_bam(raw, lay)
    with pytest.raises(AmpError):
        eng.decoded_write_bam(head)                                   # decoded but not trimmed
    eng.process_decoded(trim=False, pileup=True)
    with pytest.raises(AmpError):
        eng.decoded_write_bam(head)                                   # pileup only: still no trim outputs
    eng.process_decoded(trim=True, pileup=False, download=False)
    data, nrec = eng.decoded_write_bam(head)
    got = alnio._read_bam(data.tobytes())
    assert got.n == nrec and 0 < nrec <= b.n
    payload = np.arange(200_000, dtype=np.uint8)
    with pytest.raises(AmpError):
        eng.bgzf_deflate(payload, np.array([0, 100_000, 200_000], np.int64))          # blocks longer than 0xff00
    with pytest.raises(AmpError):
        eng.bgzf_deflate(payload, np.array([0, 50_000, 100_000], np.int64))           # table does not cover the data
    out = eng.bgzf_deflate(payload, np.array([0, 50_000, 100_000, 150_000, 200_000], np.int64))
    assert alnio.bgzf_decompress(out.tobytes()).tobytes() == payload.tobytes()


@pytest.mark.gpu
def test_device_bam_with_a_header_longer_than_a_block(tmp_path, monkeypatch):
    """4,000 reference sequences: the BAM header spans two BGZF blocks, the first record starts in the middle of one.  Device
    decode == host decode, and the device writer's file (header cut across blocks again) reads back with the same records."""
    import os
    from amplipy_b200 import alnio, cli
    g, prim, amps = _scheme(L=6000, n_amp=18)
    b = synth.illumina_batch(g, amps, 20_000, seed=81)
    refs = [("ref", 6000)] + [("decoy_%05d_with_a_long_name" % k, 1000 + k) for k in range(4000)]
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % r for r in refs) + "@PG\tID:synth\tPN:synth\n"
    d = str(tmp_path)
    j = lambda n: os.path.join(d, n)
    alnio.write_bam(j("in.bam"), text, refs, b)
    raw = open(j("in.bam"), "rb").read()
    lay = alnio.bam_layout(raw)
    assert lay["body_off"] > 0xff00 and len(lay["refs"]) == 4001
    a = alnio._read_bam(raw)
    eng = make_engine(ref_len=6000)
    info = eng.decode_bam(raw, lay)
    got, rec_off = eng.decoded_batch()
    assert info["n"] == b.n
    for f in ("pos", "flag", "tlen", "cig_off", "cigar", "seq_off", "seq", "qual_off", "qual"):
        assert np.array_equal(getattr(got, f), getattr(a.batch, f)), f
    assert np.array_equal(rec_off.astype(np.int64), a.bam_rec_off)
    with open(j("p.bed"), "w") as f:
        for k, (s, e) in enumerate(prim):
            f.write("ref\t%d\t%d\tp%d\n" % (s, e, k))
    with open(j("ref.fas"), "w") as f:
        f.write(">ref\n%s\n" % g)
    monkeypatch.delenv("AMPLIPY_BAM_LEVEL", raising=False)
    cli.main(["trim", "-i", j("in.bam"), "-p", j("p.bed"), "-r", j("ref.fas"), "-o", j("dev.bam")])
    monkeypatch.setenv("AMPLIPY_BAM_LEVEL", "6")
    cli.main(["trim", "-i", j("in.bam"), "-p", j("p.bed"), "-r", j("ref.fas"), "-o", j("zlib.bam")])
    x, z = alnio.read_alignments(j("dev.bam")), alnio.read_alignments(j("zlib.bam"))
    assert x.n == z.n and x.n > 0 and x.refs == z.refs and len(x.refs) == 4001
    for f in ("pos", "flag", "tlen", "cig_off", "cigar", "seq_off", "seq", "qual_off", "qual"):
        assert np.array_equal(getattr(x.batch, f), getattr(z.batch, f)), f
    # and the device decoder reads the device writer's file
    out = open(j("dev.bam"), "rb").read()
    info2 = eng.decode_bam(out, alnio.bam_layout(out))
    assert info2["n"] == x.n
