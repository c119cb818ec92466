"""One end-to-end step from BAM bytes (decode on the device -> fused kernel -> call), repeated; prints wall times.
usage: python profiles/tools/bam_step.py [reads] [repeats]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import bench
from amplipy_b200 import alnio
from amplipy_b200.engine import Engine
from amplipy_b200.primers import find_overlapping_primers, max_primer_len
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
g, prim, b = bench.make_workload(n, 2)
tables = find_overlapping_primers(len(g), prim, 0)
eng = Engine(ref_len=len(g), primer_tables=tables, max_primer_len=max_primer_len(prim))
eng.set_reference(g)
import tempfile
with tempfile.TemporaryDirectory() as td:
    p = os.path.join(td, "in.bam")
    alnio.write_bam(p, "@HD\tVN:1.6\n@SQ\tSN:ref\tLN:%d\n@PG\tID:x\tPN:x\n" % len(g), [("ref", len(g))], b, level=6)
    raw = np.fromfile(p, np.uint8)
lay = alnio.bam_layout(raw.tobytes())
pin = torch.from_numpy(raw).pin_memory().numpy()
out = None
for r in range(reps):
    eng.reset()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    info = eng.decode_bam(pin, lay)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if out is None:
        out = tuple(torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a.view(np.int32) if a.dtype == np.uint32 else a).pin_memory().numpy().view(a.dtype) for a in eng.alloc_decoded_trim_out())
    out = eng.process_decoded(out=out)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    res = eng.call(None, pinned=True)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print("decode %.2f ms  process %.2f ms  call %.2f ms  (bam %d bytes, %d blocks, %d reads)" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, raw.size, lay["in_off"].size, info["n"]))
