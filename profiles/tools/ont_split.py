"""Experiment: the ONT kernel's trim-only and pileup-only instances timed apart, next to the fused one (is a two-kernel split,
each with half the code, faster than one kernel that misses in the instruction cache?).  usage: ont_split.py [reads]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from amplipy_b200.engine import Engine
from amplipy_b200.primers import find_overlapping_primers, max_primer_len

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
g, prim, b = bench.make_workload(n, 4, "ont")
tables = find_overlapping_primers(bench.L_GENOME, prim, 0)
eng = Engine(ref_len=bench.L_GENOME, primer_tables=tables, max_primer_len=max_primer_len(prim), device=0, ins_slots=1 << 24, ins_arena_bytes=1 << 30)
eng.set_reference(g)
d = eng.upload(b)
eng.reserve(b.n, int(b.cig_off[-1]))
st = torch.cuda.Stream()
def run(trim, pile, reps=8):
    with torch.cuda.stream(st):
        for _ in range(3):
            eng.reset_async(st.cuda_stream); eng.process_device(d, trim=trim, pileup=pile, stream=st.cuda_stream)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            eng.reset_async(st.cuda_stream)
            a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); eng.process_device(d, trim=trim, pileup=pile, stream=st.cuda_stream); c.record(st)
            torch.cuda.synchronize(); ts.append(a.elapsed_time(c))
    return float(np.median(ts))
print("reads %d  fused %.3f ms  trim-only %.3f ms  pileup-only (untrimmed CIGARs) %.3f ms  err %d" % (n, run(True, True), run(True, False), run(False, True), eng.error_flags()))
