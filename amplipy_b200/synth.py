"""Seeded synthetic amplicon data -- numpy-vectorised so that 1M+ reads
are generated in seconds on the GPU box (no network, no real BAMs there).

All generated reads are *reference-legal* (SURVEY.md H6): 4-column BED, bases in ACGTN, no read
ending in an insertion, every aligned base inside the genome.

Shapes:
  * ``illumina_batch``  2x150 paired reads over an ARTIC-like tiling scheme (configs 1,2,3,5)
  * ``ont_batch``       single-end ~400 bp reads with a high indel rate (config 4)
  * ``fuzz_records``    small adversarial CIGAR/quality cases for parity tests
"""
import numpy as np

from .batch import ReadBatch

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
# ASCII -> BAM nibble
_NIB = np.full(256, 15, np.uint8)
for _i, _c in enumerate(b"=ACMGRSVTWYHKDBN"):
    _NIB[_c] = _i


def random_genome(L, seed=0):
    rng = np.random.default_rng(seed)
    return _ACGT[rng.integers(0, 4, L)].tobytes().decode()


def make_scheme(L, n_amplicons=98, amp_len=400, primer_len=(22, 30), seed=0, n_alt=0):
    """ARTIC-like tiling scheme: returns (primers, amplicons).

    primers   : list of (start, end, name) half-open, sorted -- one fwd + one rev per amplicon,
                plus ``n_alt`` shifted "alt" primers (v4.1-like overlapping primer records).
    amplicons : int32 array [n, 2] of (start, end) -- fragment span incl. primers.
    Real ARTIC coordinates are not available offline; this is a generated stand-in.
    """
    rng = np.random.default_rng(seed)
    first = 30
    step = (L - 60 - amp_len) / max(n_amplicons - 1, 1)
    primers = []
    amps = np.zeros((n_amplicons, 2), np.int32)
    for i in range(n_amplicons):
        a = int(round(first + i * step)) + int(rng.integers(-5, 6))
        a = max(a, 1)
        b = min(a + amp_len + int(rng.integers(-15, 16)), L - 5)
        lf = int(rng.integers(primer_len[0], primer_len[1] + 1))
        lr = int(rng.integers(primer_len[0], primer_len[1] + 1))
        amps[i] = (a, b)
        primers.append((a, a + lf, "amp_%d_LEFT" % (i + 1)))
        primers.append((b - lr, b, "amp_%d_RIGHT" % (i + 1)))
    for k in range(n_alt):
        i = int(rng.integers(0, n_amplicons))
        a, b = int(amps[i, 0]), int(amps[i, 1])
        sh = int(rng.integers(3, 12))
        if k & 1:
            primers.append((a + sh, a + sh + int(rng.integers(primer_len[0], primer_len[1] + 1)), "amp_%d_LEFT_alt" % (i + 1)))
        else:
            primers.append((b - sh - int(rng.integers(primer_len[0], primer_len[1] + 1)), b - sh, "amp_%d_RIGHT_alt" % (i + 1)))
    primers.sort()
    return primers, amps


def write_bed(path, primers, rname="ref"):
    with open(path, "w") as f:
        for s, e, name in primers:
            f.write("%s\t%d\t%d\t%s\n" % (rname, s, e, name))


def write_fasta(path, name, seq, width=70):
    with open(path, "w") as f:
        f.write(">%s\n" % name)
        for i in range(0, len(seq), width):
            f.write(seq[i:i + width] + "\n")


def _quals(rng, shape, probs=(0.70, 0.20, 0.08, 0.02), levels=(37, 25, 11, 2)):
    u = rng.random(shape, dtype=np.float32)
    q = np.full(shape, levels[-1], np.uint8)
    acc = 0.0
    for p, lv in zip(probs[:-1], levels[:-1]):
        q[(u >= acc) & (u < acc + p)] = lv
        acc += p
    return q


def _pack_rows(nib, lens):
    """nib: u8[n, lmax] nibble codes (garbage beyond lens). -> (packed bytes flat, byte offsets)."""
    n, lmax = nib.shape
    if lmax & 1:
        nib = np.concatenate([nib, np.zeros((n, 1), np.uint8)], axis=1)
        lmax += 1
    col = np.arange(lmax, dtype=np.int32)[None, :]
    nib = np.where(col < lens[:, None], nib, 0).astype(np.uint8)
    packed = (nib[:, 0::2] << 4) | nib[:, 1::2]
    nb = (lens.astype(np.int64) + 1) // 2
    keep = np.arange(lmax // 2, dtype=np.int64)[None, :] < nb[:, None]
    off = np.zeros(n + 1, np.int64)
    np.cumsum(nb, out=off[1:])
    return packed[keep], off


def _ragged(mat, lens):
    n, lmax = mat.shape
    keep = np.arange(lmax, dtype=np.int64)[None, :] < lens[:, None]
    off = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    return mat[keep], off


def _illumina_chunk(rng, ref, amps, n_pairs, read_len, sub_rate, p_ins, p_del, p_clip, p_short, p_hard,
                    snv_pos, snv_alt, snv_af, amp_weights):
    L = ref.shape[0]
    a_idx = rng.choice(amps.shape[0], size=n_pairs, p=amp_weights)
    A = amps[a_idx, 0].astype(np.int64)
    B = amps[a_idx, 1].astype(np.int64)
    short = rng.random(n_pairs) < p_short
    ins_sz = np.where(short, rng.integers(60, 200, n_pairs), B - A)
    left_anchor = rng.random(n_pairs) < 0.5
    fA = np.where(short & ~left_anchor, B - ins_sz, A)
    fB = np.where(short & left_anchor, A + ins_sz, B)
    carries = rng.random((n_pairs, max(len(snv_pos), 1))) < (np.asarray(snv_af, np.float64)[None, :] if len(snv_pos) else 0.0)

    n = 2 * n_pairs
    is_rev = np.zeros(n, bool)
    is_rev[1::2] = True
    fA2 = np.repeat(fA, 2)
    fB2 = np.repeat(fB, 2)
    carries2 = np.repeat(carries, 2, axis=0)
    insert = fB2 - fA2
    l = np.minimum(read_len, insert).astype(np.int64)
    # clips / indel layout: [H?] S1 M1 (I|D) M2 S2 [H?]
    s1 = np.where(rng.random(n) < p_clip, rng.integers(1, 11, n), 0)
    s2 = np.where(rng.random(n) < p_clip, rng.integers(1, 11, n), 0)
    u = rng.random(n)
    has_ins = u < p_ins
    has_del = (u >= p_ins) & (u < p_ins + p_del)
    ilen = np.where(has_ins, rng.integers(1, 4, n), 0)
    dlen = np.where(has_del, rng.integers(1, 6, n), 0)
    aligned_q = l - s1 - s2 - ilen
    bad = aligned_q < 20
    s1[bad] = 0
    s2[bad] = 0
    ilen[bad] = 0
    dlen[bad] = 0
    has_ins &= ~bad
    has_del &= ~bad
    aligned_q = l - s1 - s2 - ilen
    split = np.where(has_ins | has_del, (5 + rng.random(n) * (aligned_q - 10)).astype(np.int64), aligned_q)
    m1 = split
    m2 = aligned_q - split
    refspan = m1 + dlen + m2
    pos = np.where(is_rev, fB2 - refspan, fA2)
    jit = np.where(rng.random(n) < 0.05, rng.integers(0, 4, n), 0)
    pos = np.where(is_rev, pos - jit, pos + jit)
    pos = np.clip(pos, 0, L - refspan)
    hard = rng.random(n) < p_hard
    hlen = np.where(hard, rng.integers(5, 80, n), 0)

    first_in_pair = np.zeros(n, bool)
    first_in_pair[0::2] = True
    swap = np.repeat(rng.random(n_pairs) < 0.5, 2)      # which mate is read1
    r1 = first_in_pair ^ swap
    flag = (1 + 2 + np.where(is_rev, 16, 32) + np.where(r1, 64, 128) + np.where(hard, 2048, 0)).astype(np.uint16)
    tlen = np.where(is_rev, -insert, insert).astype(np.int32)

    # query index -> reference index (int32 matrices; sparse events are drawn by index, not by full random matrices)
    lmax = int(l.max())
    j = np.arange(lmax, dtype=np.int32)[None, :]
    s1c, m1c, ilc, m2c, dlc, posc = (x.astype(np.int32)[:, None] for x in (s1, m1, ilen, m2, dlen, pos))
    jj = j - s1c
    in_m1 = (jj >= 0) & (jj < m1c)
    jj2 = jj - m1c - ilc
    in_m2 = (jj2 >= 0) & (jj2 < m2c)
    R = np.where(in_m1, posc + jj, np.where(in_m2, posc + m1c + dlc + jj2, -1))
    rnd = _ACGT[rng.integers(0, 4, (n, lmax), dtype=np.uint8)]
    base = np.where(R >= 0, ref[np.clip(R, 0, L - 1)], rnd)
    for k in range(len(snv_pos)):
        cols = snv_pos[k] - pos                       # query column of the SNV if inside M1 (no indel before it)
        rows = np.flatnonzero(carries2[:, k])
        m = (R[rows] == snv_pos[k])
        rr, cc = np.nonzero(m)
        base[rows[rr], cc] = snv_alt[k]
    tot = n * lmax
    flat = base.reshape(-1)
    ks = rng.binomial(tot, sub_rate)
    idx = rng.integers(0, tot, ks)
    flat[idx] = _ACGT[rng.integers(0, 4, ks, dtype=np.uint8)]
    kn = rng.binomial(tot, 0.001)
    idxn = rng.integers(0, tot, kn)
    flat[idxn] = ord("N")
    base = flat.reshape(n, lmax)

    lut = np.empty(256, np.uint8)
    lut[:179] = 37; lut[179:230] = 25; lut[230:250] = 11; lut[250:] = 2     # ~70 / 20 / 8 / 2 %
    qual = lut[rng.integers(0, 256, (n, lmax), dtype=np.uint8)]
    qual.reshape(-1)[idxn] = 2
    tail = rng.random(n) < 0.12
    tl = np.where(tail, rng.integers(5, 41, n), 0).astype(np.int32)
    # 3' end is the right end for forward reads, the left end for reverse reads
    rows = np.flatnonzero(tail)
    lt = l.astype(np.int32)
    in_tail = np.where(is_rev[rows, None], j < tl[rows, None], j >= (lt[rows] - tl[rows])[:, None])
    lowq = rng.integers(2, 12, in_tail.shape, dtype=np.uint8)
    qual[rows] = np.where(in_tail, lowq, qual[rows])

    # CIGAR columns (len, op); zero-length dropped
    ops = np.array([5, 4, 0, 1, 0, 4, 5], np.uint32)
    lens = np.stack([np.zeros(n, np.int64), s1, m1, np.where(has_ins, ilen, dlen), m2, s2, hlen], axis=1)
    opm = np.broadcast_to(ops[None, :], lens.shape).copy()
    opm[:, 3] = np.where(has_ins, 1, 2)
    return pos.astype(np.int32), flag, tlen, lens, opm, base, qual, l


def illumina_batch(ref_seq, amps, n_reads, seed=1, read_len=150, sub_rate=0.005, p_ins=0.02, p_del=0.02,
                   p_clip=0.03, p_short=0.08, p_hard=0.003, snvs=(), amp_weights=None, chunk_pairs=200_000,
                   sort=True):
    """Paired 2x``read_len`` reads over the amplicon scheme.  ``snvs`` = [(pos0, 'A', af), ...]."""
    rng = np.random.default_rng(seed)
    ref = np.frombuffer(ref_seq.encode(), dtype=np.uint8)
    n_pairs = (n_reads + 1) // 2
    snv_pos = [int(s[0]) for s in snvs]
    snv_alt = [ord(s[1]) for s in snvs]
    snv_af = [float(s[2]) for s in snvs]
    if amp_weights is None:
        amp_weights = np.full(amps.shape[0], 1.0 / amps.shape[0])
    parts = []
    done = 0
    while done < n_pairs:
        m = min(chunk_pairs, n_pairs - done)
        parts.append(_illumina_chunk(rng, ref, amps, m, read_len, sub_rate, p_ins, p_del, p_clip, p_short, p_hard,
                                     snv_pos, snv_alt, snv_af, amp_weights))
        done += m
    lmax = max(p[5].shape[1] for p in parts)

    def padc(x):
        return x if x.shape[1] == lmax else np.pad(x, ((0, 0), (0, lmax - x.shape[1])))
    pos = np.concatenate([p[0] for p in parts])[:n_reads]
    flag = np.concatenate([p[1] for p in parts])[:n_reads]
    tlen = np.concatenate([p[2] for p in parts])[:n_reads]
    lens = np.concatenate([p[3] for p in parts])[:n_reads]
    opm = np.concatenate([p[4] for p in parts])[:n_reads]
    base = np.concatenate([padc(p[5]) for p in parts])[:n_reads]
    qual = np.concatenate([padc(p[6]) for p in parts])[:n_reads]
    l = np.concatenate([p[7] for p in parts])[:n_reads]
    return _assemble(pos, flag, tlen, lens, opm, base, qual, l, sort)


def _assemble(pos, flag, tlen, lens, opm, base, qual, l, sort):
    if sort:
        order = np.argsort(pos, kind="stable")
        pos, flag, tlen, lens, opm, base, qual, l = (x[order] for x in (pos, flag, tlen, lens, opm, base, qual, l))
    keep = lens > 0
    cig = ((lens.astype(np.uint32) << 4) | opm)[keep].astype(np.uint32)
    cig_off = np.zeros(pos.shape[0] + 1, np.int64)
    np.cumsum(keep.sum(axis=1), out=cig_off[1:])
    seq, seq_off = _pack_rows(_NIB[base], l)
    q, qual_off = _ragged(qual, l)
    return ReadBatch(np.ascontiguousarray(pos, np.int32), np.ascontiguousarray(flag, np.uint16),
                     np.ascontiguousarray(tlen, np.int32), cig_off.astype(np.uint32), cig,
                     seq_off.astype(np.uint32), np.ascontiguousarray(seq), qual_off.astype(np.uint32),
                     np.ascontiguousarray(q)).validate()


def ont_batch(ref_seq, amps, n_reads, seed=4, sub_rate=0.02, ins_rate=0.03, del_rate=0.03, p_clip=0.3,
              chunk=50_000, sort=True):
    """Single-end ONT-like reads covering whole amplicons; many short indels (homopolymer-biased)."""
    rng = np.random.default_rng(seed)
    ref = np.frombuffer(ref_seq.encode(), dtype=np.uint8)
    outs = []
    done = 0
    while done < n_reads:
        m = min(chunk, n_reads - done)
        outs.append(_ont_chunk(rng, ref, amps, m, sub_rate, ins_rate, del_rate, p_clip))
        done += m
    b = ReadBatch.concat(outs) if len(outs) > 1 else outs[0]
    if sort:
        order = np.argsort(b.pos, kind="stable")
        b = _reorder(b, order)
    return b.validate()


def _reorder(b, order):
    """Vectorised gather of ragged reads by index array."""
    def gather(off, data):
        off = off.astype(np.int64)
        ln = np.diff(off)[order]
        new_off = np.zeros(len(order) + 1, np.int64)
        np.cumsum(ln, out=new_off[1:])
        idx = np.repeat(off[:-1][order] - new_off[:-1], ln) + np.arange(int(new_off[-1]), dtype=np.int64)
        return new_off.astype(np.uint32), np.ascontiguousarray(data[idx])
    co, c = gather(b.cig_off, b.cigar)
    so, s = gather(b.seq_off, b.seq)
    qo, q = gather(b.qual_off, b.qual)
    return ReadBatch(np.ascontiguousarray(b.pos[order]), np.ascontiguousarray(b.flag[order]),
                     np.ascontiguousarray(b.tlen[order]), co, c, so, s, qo, q)


def _ont_chunk(rng, ref, amps, n, sub_rate, ins_rate, del_rate, p_clip):
    L = ref.shape[0]
    a_idx = rng.integers(0, amps.shape[0], n)
    A = amps[a_idx, 0].astype(np.int64)
    B = amps[a_idx, 1].astype(np.int64)
    A = A + np.where(rng.random(n) < 0.2, rng.integers(0, 8, n), 0)
    B = B - np.where(rng.random(n) < 0.2, rng.integers(0, 8, n), 0)
    span = (B - A)
    smax = int(span.max())
    col = np.arange(smax, dtype=np.int64)[None, :]
    valid = col < span[:, None]
    interior = (col >= 3) & (col < (span[:, None] - 3))
    # per reference column: number of inserted bases before it, deleted or not
    k = np.where(rng.random((n, smax), dtype=np.float32) < ins_rate, rng.geometric(0.6, (n, smax)), 0)
    k = np.where(valid & interior, k, 0).astype(np.int64)
    dstart = (rng.random((n, smax), dtype=np.float32) < del_rate) & interior
    dlen = np.where(dstart, rng.geometric(0.6, (n, smax)), 0)
    # spread deletion lengths to the right
    dele = np.zeros((n, smax), bool)
    for t in range(4):
        sh = np.zeros_like(dele)
        if t == 0:
            sh = dlen > 0
        else:
            sh[:, t:] = dlen[:, :-t] > t
        dele |= sh
    dele &= valid & interior
    # flatten into alignment columns: k inserted columns then 1 ref column (M or D)
    cnt = np.where(valid, k + 1, 0)
    tot_per_read = cnt.sum(axis=1)
    flat_cnt = cnt.ravel()
    nz = flat_cnt > 0
    cell = np.flatnonzero(nz)
    reps = flat_cnt[nz]
    ends = np.cumsum(reps) - 1                       # index of each cell's ref column
    total = int(reps.sum())
    typ = np.ones(total, np.uint8)                   # 1 = I
    cell_read = cell // smax
    cell_col = cell % smax
    typ[ends] = np.where(dele.ravel()[cell], 2, 0)   # 2 = D, 0 = M
    col_read = np.repeat(cell_read, reps)
    col_ref = np.repeat(A[cell_read] + cell_col, reps)
    # soft clips
    s1 = np.where(rng.random(n) < p_clip, rng.integers(1, 31, n), 0)
    s2 = np.where(rng.random(n) < p_clip, rng.integers(1, 31, n), 0)
    # bases for query-consuming columns
    is_q = typ != 2
    refb = ref[np.clip(col_ref, 0, L - 1)]
    rnd = _ACGT[rng.integers(0, 4, total)]
    prevb = np.concatenate([[ord("A")], refb[:-1]])
    insb = np.where(rng.random(total) < 0.5, prevb, rnd)          # homopolymer-biased insertions
    b = np.where(typ == 1, insb, np.where(rng.random(total) < sub_rate, rnd, refb)).astype(np.uint8)
    # per-read quality level + noise, inserted bases a little worse, occasional low-quality stretches
    level = rng.integers(22, 33, n)
    qv = (level[col_read] + rng.integers(-6, 7, total))
    qv = np.where(typ == 1, qv - rng.integers(0, 12, total), qv)
    qv = np.where(rng.random(total) < 0.03, rng.integers(2, 15, total), qv)
    bad_read = rng.random(n) < 0.3
    bad_at = (rng.random(n) * tot_per_read).astype(np.int64)
    bad_len = rng.integers(5, 40, n)
    within_col = np.arange(total) - np.repeat(np.concatenate([[0], np.cumsum(tot_per_read)[:-1]]), tot_per_read)
    in_bad = bad_read[col_read] & (within_col >= bad_at[col_read]) & (within_col < (bad_at + bad_len)[col_read])
    qv = np.where(in_bad, rng.integers(3, 13, total), qv)
    qv = np.clip(qv, 2, 40).astype(np.uint8)
    # run-length encode column types per read
    read_start = np.zeros(n + 1, np.int64)
    np.cumsum(tot_per_read, out=read_start[1:])
    brk = np.ones(total, bool)
    brk[1:] = (typ[1:] != typ[:-1]) | (col_read[1:] != col_read[:-1])
    run_start = np.flatnonzero(brk)
    run_len = np.diff(np.concatenate([run_start, [total]]))
    run_op = typ[run_start].astype(np.uint32)
    run_read = col_read[run_start]
    runs_per_read = np.bincount(run_read, minlength=n)
    # assemble cigar with optional clips: per read [S1] runs [S2]
    has1 = s1 > 0
    has2 = s2 > 0
    ncig = runs_per_read + has1 + has2
    cig_off = np.zeros(n + 1, np.int64)
    np.cumsum(ncig, out=cig_off[1:])
    cig = np.zeros(int(cig_off[-1]), np.uint32)
    run_first = np.zeros(n + 1, np.int64)
    np.cumsum(runs_per_read, out=run_first[1:])
    dst = cig_off[:-1][run_read] + has1[run_read] + (np.arange(len(run_start)) - run_first[:-1][run_read])
    cig[dst] = (run_len.astype(np.uint32) << 4) | run_op
    cig[cig_off[:-1][has1]] = (s1[has1].astype(np.uint32) << 4) | 4
    cig[cig_off[1:][has2] - 1] = (s2[has2].astype(np.uint32) << 4) | 4
    # query sequence / quals: S1 random + query columns + S2 random
    qcols = np.flatnonzero(is_q)
    q_per_read = np.bincount(col_read[qcols], minlength=n)
    l = q_per_read + s1 + s2
    qoff = np.zeros(n + 1, np.int64)
    np.cumsum(l, out=qoff[1:])
    totq = int(qoff[-1])
    bases = _ACGT[rng.integers(0, 4, totq)].copy()
    quals = rng.integers(8, 31, totq).astype(np.uint8)
    qfirst = np.zeros(n + 1, np.int64)
    np.cumsum(q_per_read, out=qfirst[1:])
    rr = col_read[qcols]
    d2 = qoff[:-1][rr] + s1[rr] + (np.arange(len(qcols)) - qfirst[:-1][rr])
    bases[d2] = b[qcols]
    quals[d2] = qv[qcols]
    # pack nibbles per read (ragged): pad each read to even length in a flat array
    nb = (l + 1) // 2
    soff = np.zeros(n + 1, np.int64)
    np.cumsum(nb, out=soff[1:])
    nibflat = np.zeros(int(soff[-1]) * 2, np.uint8)
    ridx = np.repeat(np.arange(n), l)
    within = np.arange(totq) - qoff[:-1][ridx]
    nibflat[2 * soff[:-1][ridx] + within] = _NIB[bases]
    seq = (nibflat[0::2] << 4) | nibflat[1::2]
    flag = np.where(rng.random(n) < 0.5, 16, 0).astype(np.uint16)
    return ReadBatch(A.astype(np.int32), flag, np.zeros(n, np.int32), cig_off.astype(np.uint32), cig,
                     soff.astype(np.uint32), seq.astype(np.uint8), qoff.astype(np.uint32), quals)


# ---------------------------------------------------------------------------------------------
# adversarial small cases (python loops; parity tests only)
# ---------------------------------------------------------------------------------------------
def fuzz_records(L, n, seed=0, max_len=60, ont_like=False, edges=0.0):
    """Random well-formed reads with H/S/M/I/D/N/=/X shapes, I->D adjacency, all flag/tlen combos.
    ``edges`` = fraction of the reads placed at position 0 or ending exactly on the last reference base."""
    rng = np.random.default_rng(seed)
    recs = []
    while len(recs) < n:
        ops = []
        if rng.random() < 0.1:
            ops.append((5, int(rng.integers(1, 20))))
        if rng.random() < 0.3:
            ops.append((4, int(rng.integers(1, 12))))
        nblocks = int(rng.integers(1, 12 if ont_like else 4))
        for bi in range(nblocks):
            ops.append((int(rng.choice([0, 0, 0, 7, 8])), int(rng.integers(1, max_len if not ont_like else 25))))
            if bi != nblocks - 1:
                r = rng.random()
                if r < 0.35:
                    ops.append((1, int(rng.integers(1, 5))))
                    if rng.random() < 0.15:
                        ops.append((2, int(rng.integers(1, 4))))     # I followed by D
                elif r < 0.7:
                    ops.append((2, int(rng.integers(1, 6))))
                elif r < 0.8:
                    ops.append((3, int(rng.integers(1, 30))))
        if rng.random() < 0.08:
            ops.append((1, int(rng.integers(1, 4))))                 # insertion run into a trailing clip
            ops.append((4, int(rng.integers(1, 8))))
        elif rng.random() < 0.3:
            ops.append((4, int(rng.integers(1, 12))))
        if rng.random() < 0.1:
            ops.append((5, int(rng.integers(1, 20))))
        if rng.random() < 0.04 and ops[0][0] in (0, 7, 8):
            ops.insert(0, (1, int(rng.integers(1, 3))))              # leading insertion (q0 == 0 quirk)
        # merge equal adjacent ops (aligners never emit them)
        merged = []
        for op, ln in ops:
            if merged and merged[-1][0] == op:
                merged[-1] = (op, merged[-1][1] + ln)
            else:
                merged.append((op, ln))
        ops = merged
        qlen = sum(ln for op, ln in ops if op in (0, 1, 4, 7, 8))
        rlen = sum(ln for op, ln in ops if op in (0, 2, 3, 7, 8))
        if rlen + 2 >= L:
            continue
        pos = int(rng.integers(1, L - rlen - 1))
        if edges and rng.random() < edges:
            pos = 0 if rng.random() < 0.5 else L - rlen
        seq = "".join(rng.choice(list("ACGTN"), p=[0.24, 0.24, 0.24, 0.24, 0.04]) for _ in range(qlen))
        mode = rng.random()
        if mode < 0.5:
            qual = [int(x) for x in rng.choice([37, 25, 11, 2], p=[0.6, 0.2, 0.12, 0.08], size=qlen)]
        elif mode < 0.75:
            qual = [int(x) for x in rng.integers(0, 42, qlen)]
        else:
            k = int(rng.integers(0, qlen + 1))
            lo = [int(x) for x in rng.integers(0, 15, qlen)]
            hi = [int(x) for x in rng.integers(25, 41, qlen)]
            qual = (lo[:k] + hi[k:]) if rng.random() < 0.5 else (hi[:k] + lo[k:])
        flag = int(rng.choice([0, 16, 99, 147, 83, 163, 1, 17, 2048 + 16, 1024 + 99]))
        tl = int(rng.choice([0, qlen, qlen + 10, qlen + 25, qlen + 45, 400, 1000]))
        tlen = -tl if (flag & 16) else tl
        recs.append((pos, flag, tlen, ops, seq, qual))
    return recs
