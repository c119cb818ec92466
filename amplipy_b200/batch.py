"""Flat struct-of-arrays read batches: the host-side data layout handed to the C-ABI.

One batch = N aligned reads in input order.  All arrays are C-contiguous numpy arrays:

    pos      i32[N]     0-based leftmost reference coordinate (BAM ``pos``)
    flag     u16[N]     BAM flag word
    tlen     i32[N]     BAM template length
    cig_off  u32[N+1]   prefix offsets (in ops) into ``cigar``; n_cigar = diff
    cigar    u32[sumC]  BAM packed ops: ``len << 4 | op`` (op codes ``MIDNSHP=X`` = 0..8)
    seq_off  u32[N+1]   prefix offsets (in bytes) into ``seq``; each read occupies (l_seq+1)//2 bytes
    seq      u8[...]    BAM 4-bit bases, high nibble first, code table ``=ACMGRSVTWYHKDBN``
    qual_off u32[N+1]   prefix offsets (in bytes) into ``qual``; l_seq = diff
    qual     u8[sumL]   phred qualities (not +33)

This replaces the per-record ``pysam.AlignedSegment`` attribute reads of the reference
(``/root/reference/AmpliPy.py:450-452, 561, 700-706``): the fields above are exactly what
``trim_read`` / ``update_base_counts`` consume.
"""
from dataclasses import dataclass

import numpy as np

CIGAR_CHARS = "MIDNSHP=XB"
NIBBLE_CHARS = "=ACMGRSVTWYHKDBN"
_NIB_OF = np.full(256, 15, dtype=np.uint8)
for _i, _c in enumerate(NIBBLE_CHARS):
    _NIB_OF[ord(_c)] = _i
    _NIB_OF[ord(_c.lower())] = _i
_CHAR_OF_NIB = np.frombuffer(NIBBLE_CHARS.encode(), dtype=np.uint8)


def parse_cigar(s):
    """'31S120M' -> [(4,31),(0,120)]; '*' -> []"""
    if s == "*" or not s:
        return []
    out = []
    n = 0
    for ch in s:
        if "0" <= ch <= "9":
            n = n * 10 + ord(ch) - 48
        else:
            out.append((CIGAR_CHARS.index(ch), n))
            n = 0
    return out


def cigar_string(ops):
    if len(ops) == 0:
        return "*"
    return "".join("%d%s" % (n, CIGAR_CHARS[op]) for op, n in ops)


def pack_seq(seq_str):
    """ASCII bases -> BAM 4-bit packed bytes (high nibble first)."""
    a = _NIB_OF[np.frombuffer(seq_str.encode(), dtype=np.uint8)]
    if len(a) & 1:
        a = np.concatenate([a, np.zeros(1, np.uint8)])
    return ((a[0::2] << 4) | a[1::2]).astype(np.uint8)


def unpack_seq(packed, l_seq):
    hi = packed >> 4
    lo = packed & 15
    nib = np.empty(len(packed) * 2, np.uint8)
    nib[0::2] = hi
    nib[1::2] = lo
    return _CHAR_OF_NIB[nib[:l_seq]].tobytes().decode()


@dataclass
class ReadBatch:
    pos: np.ndarray
    flag: np.ndarray
    tlen: np.ndarray
    cig_off: np.ndarray
    cigar: np.ndarray
    seq_off: np.ndarray
    seq: np.ndarray
    qual_off: np.ndarray
    qual: np.ndarray

    @property
    def n(self):
        return int(self.pos.shape[0])

    def __len__(self):
        return self.n

    @property
    def l_seq(self):
        return np.diff(self.qual_off.astype(np.int64)).astype(np.int32)

    @property
    def n_cigar(self):
        return np.diff(self.cig_off.astype(np.int64)).astype(np.int32)

    def validate(self):
        n = self.n
        assert self.pos.dtype == np.int32 and self.flag.dtype == np.uint16 and self.tlen.dtype == np.int32
        assert self.cig_off.dtype == np.uint32 and self.cig_off.shape == (n + 1,)
        assert self.seq_off.dtype == np.uint32 and self.seq_off.shape == (n + 1,)
        assert self.qual_off.dtype == np.uint32 and self.qual_off.shape == (n + 1,)
        assert self.cigar.dtype == np.uint32 and self.seq.dtype == np.uint8 and self.qual.dtype == np.uint8
        assert int(self.cig_off[-1]) == self.cigar.shape[0]
        assert int(self.seq_off[-1]) == self.seq.shape[0]
        assert int(self.qual_off[-1]) == self.qual.shape[0]
        l = np.diff(self.qual_off.astype(np.int64))
        assert np.array_equal(np.diff(self.seq_off.astype(np.int64)), (l + 1) // 2)
        for a in (self.pos, self.flag, self.tlen, self.cig_off, self.cigar, self.seq_off, self.seq,
                  self.qual_off, self.qual):
            assert a.flags["C_CONTIGUOUS"]
        return self

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_records(cls, recs):
        """recs: iterable of (pos0, flag, tlen, [(op,len)...], seq_str, qual_iterable)."""
        recs = list(recs)
        n = len(recs)
        pos = np.empty(n, np.int32)
        flag = np.empty(n, np.uint16)
        tlen = np.empty(n, np.int32)
        cig_off = np.zeros(n + 1, np.uint32)
        seq_off = np.zeros(n + 1, np.uint32)
        qual_off = np.zeros(n + 1, np.uint32)
        cig, seqs, quals = [], [], []
        for i, (p, f, t, ops, s, q) in enumerate(recs):
            pos[i] = p
            flag[i] = f
            tlen[i] = t
            cig.append(np.array([(ln << 4) | op for op, ln in ops], dtype=np.uint32))
            ps = pack_seq(s)
            seqs.append(ps)
            qa = np.asarray(list(q), dtype=np.uint8)
            assert len(qa) == len(s), "seq/qual length mismatch"
            quals.append(qa)
            cig_off[i + 1] = cig_off[i] + len(ops)
            seq_off[i + 1] = seq_off[i] + len(ps)
            qual_off[i + 1] = qual_off[i] + len(qa)
        cat = lambda xs, dt: (np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt))
        return cls(pos, flag, tlen, cig_off, cat(cig, np.uint32), seq_off, cat(seqs, np.uint8),
                   qual_off, cat(quals, np.uint8)).validate()

    @classmethod
    def from_sam_lines(cls, lines):
        recs = []
        for line in lines:
            if not line.strip() or line.startswith("@"):
                continue
            f = line.rstrip("\r\n").split("\t")
            seq = "" if f[9] == "*" else f[9]
            qual = [] if f[10] == "*" else [ord(c) - 33 for c in f[10]]
            recs.append((int(f[3]) - 1, int(f[1]), int(f[8]), parse_cigar(f[5]), seq, qual))
        return cls.from_records(recs)

    @classmethod
    def concat(cls, batches):
        batches = list(batches)
        if len(batches) == 1:
            return batches[0]

        def cat_off(name):
            outs = [np.zeros(1, np.uint64)]
            base = 0
            for b in batches:
                o = getattr(b, name).astype(np.uint64)
                outs.append(o[1:] + base)
                base += int(o[-1])
            r = np.concatenate(outs)
            assert int(r[-1]) < 2 ** 32, "batch too large for 32-bit offsets"
            return r.astype(np.uint32)
        return cls(np.concatenate([b.pos for b in batches]), np.concatenate([b.flag for b in batches]),
                   np.concatenate([b.tlen for b in batches]), cat_off("cig_off"),
                   np.concatenate([b.cigar for b in batches]), cat_off("seq_off"),
                   np.concatenate([b.seq for b in batches]), cat_off("qual_off"),
                   np.concatenate([b.qual for b in batches]))

    def slice(self, lo, hi):
        """Reads [lo, hi) as a new batch with rebased offsets (views where possible)."""
        def sub(off, data):
            a, b = int(off[lo]), int(off[hi])
            return (off[lo:hi + 1].astype(np.int64) - a).astype(np.uint32), data[a:b]
        co, c = sub(self.cig_off, self.cigar)
        so, s = sub(self.seq_off, self.seq)
        qo, q = sub(self.qual_off, self.qual)
        return ReadBatch(np.ascontiguousarray(self.pos[lo:hi]), np.ascontiguousarray(self.flag[lo:hi]),
                         np.ascontiguousarray(self.tlen[lo:hi]), co, np.ascontiguousarray(c), so,
                         np.ascontiguousarray(s), qo, np.ascontiguousarray(q))

    def take(self, idx):
        """Gather reads by index array (python-level; for tests and small inputs)."""
        return ReadBatch.from_records([self.record(int(i)) for i in idx])

    # ------------------------------------------------------------------ per-record views (tests, SAM text)
    def cigartuples(self, i):
        a, b = int(self.cig_off[i]), int(self.cig_off[i + 1])
        return [(int(c & 15), int(c >> 4)) for c in self.cigar[a:b]]

    def record(self, i):
        l = int(self.qual_off[i + 1]) - int(self.qual_off[i])
        s = unpack_seq(self.seq[int(self.seq_off[i]):int(self.seq_off[i + 1])], l)
        q = self.qual[int(self.qual_off[i]):int(self.qual_off[i + 1])]
        return (int(self.pos[i]), int(self.flag[i]), int(self.tlen[i]), self.cigartuples(i), s, q.tolist())

    def sam_lines(self, rname="ref", name_prefix="r"):
        """Minimal SAM text for each read (used to feed the reference when generating goldens)."""
        out = []
        for i in range(self.n):
            p, f, t, ops, s, q = self.record(i)
            out.append("\t".join([
                "%s%d" % (name_prefix, i), str(f), rname if not (f & 4) else "*", str(p + 1), "60",
                cigar_string(ops), "=" if (f & 1) else "*", str(max(p + 1 + t, 1)) if (f & 1) else "0", str(t),
                s if s else "*", "".join(chr(x + 33) for x in q) if q else "*"]))
        return out

    def algorithmic_bytes(self):
        """Input bytes of the hot path, each counted once."""
        return 22 * self.n + 4 * int(self.cigar.shape[0]) + int(self.seq.shape[0]) + int(self.qual.shape[0])
