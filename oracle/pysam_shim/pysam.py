"""Minimal ``pysam`` stand-in -- TEST INFRASTRUCTURE ONLY (never imported by the product).

``pysam==0.17.0`` (``/root/reference/requirements.txt:1``) is not installable in the build
container.  This module restates the small slice of its surface that
``/root/reference/AmpliPy.py`` touches, so that the *unmodified* reference can be imported
and executed here to produce golden vectors (``tests/golden/make_golden.py``).

The semantics below are restated from pysam 0.17 / htslib 1.13 behaviour and cannot be
re-verified in this container (no pysam, no htslib).  They are kept in this single file so
they can be re-validated wherever a real pysam exists:

* ``query_length``            = ``l_seq`` (0 when SEQ is ``*``)
* ``reference_end``           = ``pos + max(1, sum(len of M,D,N,=,X))`` (htslib ``bam_endpos``);
                                ``None`` if unmapped or without CIGAR
* ``reference_length``        = ``reference_end - reference_start``
* ``query_alignment_start``   = sum of leading ``S`` lengths, leading ``H`` skipped
* ``query_alignment_end``     = ``l_seq`` minus trailing ``S`` lengths, scanning ops from the
                                last one down to index 1 (op 0 is never inspected), ``H`` skipped
* ``get_aligned_pairs()``     : ``M,=,X -> (q,r)``; ``I,S,P -> (q,None)``; ``D,N -> (None,r)``; ``H`` nothing
* assigning ``cigartuples`` / ``reference_start`` rewrites only those fields

Call sites in the reference: ``AmpliPy.py:271-291,310,316-359,450-452,460-463,513-520,558-566,
591-597,625,653-658,686,700-706,896,902,910-911,947-952``.
"""
import sys
from array import array

_CIGAR_CHARS = "MIDNSHP=XB"
_CONSUME_Q = (True, True, False, False, True, False, False, True, True, False)
_CONSUME_R = (True, False, True, True, False, False, False, True, True, False)

_verbosity = 3


def set_verbosity(v):
    global _verbosity
    old = _verbosity
    _verbosity = v
    return old


def parse_cigar_string(s):
    if s == "*" or s == "":
        return None
    out = []
    n = 0
    for ch in s:
        if ch.isdigit():
            n = n * 10 + ord(ch) - 48
        else:
            out.append((_CIGAR_CHARS.index(ch), n))
            n = 0
    return out


def cigar_to_string(tuples):
    if not tuples:
        return "*"
    return "".join("%d%s" % (n, _CIGAR_CHARS[op]) for op, n in tuples)


class AlignedSegment(object):
    """Mutable record with the pysam attribute names the reference reads or writes."""

    def __init__(self):
        self.query_name = "*"
        self.flag = 0
        self.reference_name = "*"
        self.reference_start = -1          # 0-based
        self.mapping_quality = 0
        self._cigar = None                 # list of (op, len) or None
        self.next_reference_name = "*"
        self.next_reference_start = -1     # 0-based
        self.template_length = 0
        self.query_sequence = None         # str or None
        self.query_qualities = None        # array('B') or None
        self.tags_text = []                # raw SAM optional fields, passed through

    # --- construction -------------------------------------------------------------------
    @classmethod
    def from_sam_line(cls, line):
        f = line.rstrip("\r\n").split("\t")
        s = cls()
        s.query_name = f[0]
        s.flag = int(f[1])
        s.reference_name = f[2]
        s.reference_start = int(f[3]) - 1
        s.mapping_quality = int(f[4])
        s._cigar = parse_cigar_string(f[5])
        s.next_reference_name = f[6]
        s.next_reference_start = int(f[7]) - 1
        s.template_length = int(f[8])
        s.query_sequence = None if f[9] == "*" else f[9].upper()  # BAM 4-bit decode is upper-case
        s.query_qualities = None if f[10] == "*" else array("B", [ord(c) - 33 for c in f[10]])
        s.tags_text = f[11:]
        return s

    def to_sam_line(self):
        seq = "*" if self.query_sequence is None else self.query_sequence
        qual = "*" if self.query_qualities is None else "".join(chr(q + 33) for q in self.query_qualities)
        f = [self.query_name, str(self.flag), self.reference_name, str(self.reference_start + 1),
             str(self.mapping_quality), cigar_to_string(self._cigar), self.next_reference_name,
             str(self.next_reference_start + 1), str(self.template_length), seq, qual] + list(self.tags_text)
        return "\t".join(f)

    # --- flags ---------------------------------------------------------------------------
    @property
    def is_paired(self):
        return (self.flag & 1) != 0

    @property
    def is_unmapped(self):
        return (self.flag & 4) != 0

    @property
    def is_reverse(self):
        return (self.flag & 16) != 0

    # --- CIGAR ---------------------------------------------------------------------------
    @property
    def cigartuples(self):
        if not self._cigar:
            return None
        return list(self._cigar)

    @cigartuples.setter
    def cigartuples(self, values):
        self._cigar = [(int(op), int(n)) for op, n in values] if values is not None else None

    @property
    def cigarstring(self):
        return None if not self._cigar else cigar_to_string(self._cigar)

    # --- derived coordinates (htslib bam_endpos / pysam getQueryStart / getQueryEnd) ------
    @property
    def query_length(self):
        return 0 if self.query_sequence is None else len(self.query_sequence)

    @property
    def reference_end(self):
        if self.is_unmapped or not self._cigar:
            return None
        rlen = sum(n for op, n in self._cigar if _CONSUME_R[op])
        if rlen == 0:
            rlen = 1
        return self.reference_start + rlen

    @property
    def reference_length(self):
        e = self.reference_end
        return None if e is None else e - self.reference_start

    @property
    def query_alignment_start(self):
        start = 0
        for op, n in (self._cigar or ()):
            if op == 5:      # H
                continue
            elif op == 4:    # S
                start += n
            else:
                break
        return start

    @property
    def query_alignment_end(self):
        end = self.query_length
        cig = self._cigar or ()
        for k in range(len(cig) - 1, 0, -1):   # op 0 is never inspected
            op, n = cig[k]
            if op == 5:
                continue
            elif op == 4:
                end -= n
            else:
                break
        return end

    @property
    def query_alignment_qualities(self):
        if self.query_qualities is None:
            return None
        return self.query_qualities[self.query_alignment_start:self.query_alignment_end]

    def get_aligned_pairs(self):
        out = []
        q = 0
        r = self.reference_start
        for op, n in (self._cigar or ()):
            if op == 0 or op == 7 or op == 8:
                for i in range(n):
                    out.append((q + i, r + i))
                q += n
                r += n
            elif op == 1 or op == 4 or op == 6:
                for i in range(n):
                    out.append((q + i, None))
                q += n
            elif op == 2 or op == 3:
                for i in range(n):
                    out.append((None, r + i))
                r += n
            # op 5 (H): nothing
        return out


class _Header(object):
    def __init__(self, lines):
        self.lines = list(lines)

    def to_dict(self):
        d = {}
        for l in self.lines:
            if not l.startswith("@") or l.startswith("@CO"):
                continue
            f = l.rstrip("\r\n").split("\t")
            key = f[0][1:]
            rec = {}
            for kv in f[1:]:
                k, _, v = kv.partition(":")
                rec[k] = v
            if key == "HD":
                d["HD"] = rec
            else:
                d.setdefault(key, []).append(rec)
        return d


def _header_lines_from_dict(d):
    order = {"HD": ["VN", "SO", "GO", "SS"], "SQ": ["SN", "LN"], "RG": ["ID"], "PG": ["ID", "PN", "CL", "PP", "DS", "VN"]}
    out = []

    def fmt(key, rec):
        first = [k for k in order.get(key, []) if k in rec]
        rest = [k for k in rec if k not in first]
        return "@" + key + "".join("\t%s:%s" % (k, rec[k]) for k in first + rest)
    if "HD" in d:
        out.append(fmt("HD", d["HD"]))
    for key in ("SQ", "RG", "PG"):
        for rec in d.get(key, []):
            out.append(fmt(key, rec))
    return out


class AlignmentFile(object):
    """SAM-text reader/writer ('-' = stdin/stdout).  BAM is out of the shim's scope."""

    def __init__(self, fn, mode="r", header=None):
        self.filename = fn
        self.mode = mode
        self.written = []          # AlignedSegment snapshots (SAM lines), for inspection by tests
        if "b" in mode:
            raise NotImplementedError("pysam shim: BAM not supported; use .sam")
        if mode.startswith("r"):
            self._fh = sys.stdin if fn == "-" else open(fn, "r")
            hdr = []
            self._first = None
            for line in self._fh:
                if line.startswith("@"):
                    hdr.append(line.rstrip("\r\n"))
                else:
                    self._first = line
                    break
            self.header = _Header(hdr)
        else:
            self._fh = sys.stdout if fn == "-" else open(fn, "w")
            if isinstance(header, dict):
                lines = _header_lines_from_dict(header)
            elif header is None:
                lines = []
            else:
                lines = header.lines
            self.header = _Header(lines)
            for l in lines:
                self._fh.write(l + "\n")

    def __iter__(self):
        if self._first is not None:
            if self._first.strip():
                yield AlignedSegment.from_sam_line(self._first)
            self._first = None
        for line in self._fh:
            if line.strip():
                yield AlignedSegment.from_sam_line(line)

    def write(self, s):
        line = s.to_sam_line()
        self.written.append(line)
        self._fh.write(line + "\n")

    def close(self):
        if self._fh not in (sys.stdin, sys.stdout):
            self._fh.close()


# ------------------------------------------------------------------------------------------
# VCF side (AmpliPy.py:261-293, 941-952).  Records are kept structurally in ``.records`` and a
# text rendering is written; the text form of Float INFO follows htslib's "%g" of a float32.
# ------------------------------------------------------------------------------------------
class VariantHeader(object):
    def __init__(self):
        self.samples = []
        self.meta = []

    def add_sample(self, name):
        self.samples.append(name)

    def add_meta(self, key, value=None, items=None):
        if items is not None:
            body = ",".join(('%s="%s"' % (k, v)) if k == "Description" else ("%s=%s" % (k, v)) for k, v in items)
            self.meta.append("##%s=<%s>" % (key, body))
        else:
            self.meta.append("##%s=%s" % (key, value))


class _SampleFields(dict):
    pass


class VariantRecord(object):
    def __init__(self, header, contig, start, stop, alleles, info, filter):
        self.contig = contig
        self.start = start
        self.stop = stop
        self.alleles = tuple(alleles)
        self.info = dict(info)
        self.filter = filter
        self.samples = {name: _SampleFields() for name in header.samples}


def _fmt_float32(x):
    import struct
    y = struct.unpack("f", struct.pack("f", float(x)))[0]
    return "%g" % y


class VariantFile(object):
    def __init__(self, fn, mode="w", header=None):
        self.header = header
        self.records = []
        self._fh = sys.stdout if fn == "-" else open(fn, "w")
        self._fh.write("##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\">\n")
        for m in header.meta:
            self._fh.write(m + "\n")
        self._fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s\n" % "\t".join(header.samples))

    def new_record(self, contig=None, start=0, stop=0, alleles=None, info=None, filter=None):
        return VariantRecord(self.header, contig, start, stop, alleles, info, filter)

    def write(self, rec):
        self.records.append(rec)
        info = []
        for k, v in rec.info.items():
            if k == "REF_FREQ":
                info.append("%s=%s" % (k, _fmt_float32(v)))
            else:
                info.append("%s=%s" % (k, v))
        gt = rec.samples[self.header.samples[0]].get("GT", ())
        self._fh.write("\t".join([rec.contig, str(rec.start + 1), ".", rec.alleles[0], ",".join(rec.alleles[1:]),
                                  ".", rec.filter, ";".join(info), "GT", "/".join(str(g) for g in gt)]) + "\n")

    def close(self):
        if self._fh is not sys.stdout:
            self._fh.close()
