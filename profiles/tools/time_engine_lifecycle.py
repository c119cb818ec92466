import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench
from amplipy_b200.engine import Engine
from amplipy_b200.primers import find_overlapping_primers, max_primer_len
g, prim, b = bench.make_workload(1_000_000, 2, "illumina")
tables = find_overlapping_primers(bench.L_GENOME, prim, 0)
def T(label, f):
    t = time.perf_counter(); r = f(); print("%-28s %.3f s" % (label, time.perf_counter() - t)); return r
e = T("create", lambda: Engine(ref_len=bench.L_GENOME, primer_tables=tables, max_primer_len=max_primer_len(prim)))
T("destroy (fresh)", e.close)
e = T("create", lambda: Engine(ref_len=bench.L_GENOME, primer_tables=tables, max_primer_len=max_primer_len(prim)))
T("process 1M (pageable)", lambda: e.process(b))
T("process 1M again", lambda: e.process(b))
T("counts+ins+call", lambda: (e.counts(), e.insertions(), e.call(g)))
T("destroy (after use)", e.close)
