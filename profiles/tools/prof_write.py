"""Where the time of writing the trimmed BAM goes (alnio.write_alignments, BAM in -> BAM out), step by step.  usage: prof_write.py [reads]"""
import os, sys, time, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from amplipy_b200 import alnio
from amplipy_b200.engine import Engine, TrimResult
from amplipy_b200.primers import find_overlapping_primers, max_primer_len

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
g, prim, b = bench.make_workload(n, 2, "illumina")
d = tempfile.mkdtemp()
p = os.path.join(d, "in.bam")
alnio.write_bam(p, "@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:ref\tLN:%d\n@PG\tID:synth\tPN:synth\n" % bench.L_GENOME, [("ref", bench.L_GENOME)], b)
raw = open(p, "rb").read()
tables = find_overlapping_primers(bench.L_GENOME, prim, 0)
eng = Engine(ref_len=bench.L_GENOME, primer_tables=tables, max_primer_len=max_primer_len(prim), device=0)
eng.decode_bam(raw, alnio.bam_layout(raw)); outs = eng.process_decoded(trim=True, pileup=True)
aln = alnio._read_bam(raw)
trim = TrimResult(aln.batch, *outs)
lib = alnio.hostio(); _p, _ll = alnio._p, alnio._ll
for rep in range(2):
    T = []; t = time.perf_counter()
    def tick(name):
        global t
        now = time.perf_counter(); T.append((name, now - t)); t = now
    sel = np.flatnonzero(trim.keep).astype(np.int64); tick("select")
    bb = aln.batch
    size = lib.amp_bam_rewrite(_p(aln.bam_buf), _p(aln.bam_rec_off), _p(sel), _ll(len(sel)), _p(trim.pos), _p(trim.ncig), _p(bb.cig_off), _p(trim.cigar), None); tick("rewrite size")
    body = np.empty(int(size) + 8, np.uint8); tick("alloc")
    lib.amp_bam_rewrite(_p(aln.bam_buf), _p(aln.bam_rec_off), _p(sel), _ll(len(sel)), _p(trim.pos), _p(trim.ncig), _p(bb.cig_off), _p(trim.cigar), _p(body)); tick("rewrite")
    body = body[:int(size)]
    ooff = np.empty(len(sel) + 1, np.int64)
    lib.amp_bam_rewrite_offsets(_p(aln.bam_buf), _p(aln.bam_rec_off), _p(sel), _ll(len(sel)), _p(trim.ncig), _p(ooff)); tick("offsets")
    head = alnio._bam_header_bytes(aln.header_text, aln.refs); hb = np.frombuffer(head, np.uint8)
    bounds = np.unique(np.concatenate([[0], hb.size + ooff]).astype(np.int64)); tick("bounds (unique)")
    data = np.concatenate([hb, body]); tick("concatenate")
    nb = int(lib.amp_bgzf_plan(_p(bounds), _ll(bounds.size), None, _ll(0))); bstart = np.empty(nb + 1, np.int64); lib.amp_bgzf_plan(_p(bounds), _ll(bounds.size), _p(bstart), _ll(nb)); tick("plan")
    out = eng.bgzf_deflate(data, bstart); tick("device deflate (H2D + kernels + D2H)")
    with open(os.path.join(d, "out%d.bam" % rep), "wb") as f:
        f.write(out)
    tick("file write")
    print("rep", rep, "kept", len(sel), "bytes in", data.size, "out", out.size, " | ".join("%s %.1f ms" % (k, v * 1e3) for k, v in T), "| total %.1f ms" % (sum(v for _, v in T) * 1e3))
