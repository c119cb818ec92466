"""Where one e2e step (reset + amp_process_host + amp_call) spends its wall time."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from amplipy_b200.batch import ReadBatch
from amplipy_b200.engine import Engine
from amplipy_b200.primers import find_overlapping_primers, max_primer_len
g, prim, b = bench.make_workload(1_000_000, 2, "illumina")
tables = find_overlapping_primers(len(g), prim, 0)
eng = Engine(ref_len=len(g), primer_tables=tables, max_primer_len=max_primer_len(prim))
eng.set_reference(g)
def pin(a):
    if a.dtype == np.uint16: t = torch.from_numpy(np.ascontiguousarray(a.view(np.int16))).pin_memory(); return t, t.numpy().view(np.uint16)
    if a.dtype == np.uint32: t = torch.from_numpy(np.ascontiguousarray(a.view(np.int32))).pin_memory(); return t, t.numpy().view(np.uint32)
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); return t, t.numpy()
keep = []; hb = []
for f in ("pos", "flag", "tlen", "cig_off", "cigar", "seq_off", "seq", "qual_off", "qual"):
    t, v = pin(getattr(b, f)); keep.append(t); hb.append(v)
pb = ReadBatch(*hb)
pouts = []
for a in Engine.alloc_trim_out(pb):
    t, v = pin(a); keep.append(t); pouts.append(v)
pouts = tuple(pouts)
def timed(fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return (time.perf_counter() - t) * 1e3, r
for it in range(6):
    t_r, _ = timed(eng.reset)
    t_p, _ = timed(lambda: eng.process(pb, trim=True, pileup=True, out=pouts))
    t_c, _ = timed(lambda: eng.call(None, pinned=True))
    print("reset %.3f ms  process %.3f ms  call %.3f ms  sum %.3f" % (t_r, t_p, t_c, t_r + t_p + t_c), flush=True)
t = time.perf_counter()
for _ in range(10):
    eng.reset(); eng.process(pb, trim=True, pileup=True, out=pouts); eng.call(None, pinned=True)
torch.cuda.synchronize()
print("back to back: %.3f ms per step" % ((time.perf_counter() - t) * 100))
