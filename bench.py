#!/usr/bin/env python3
"""bench.py -- reads/s trimmed + piled-up (+ called) on synthetic SARS-CoV-2-shaped amplicon data.

    python bench.py --gpus N --steps K --warmup W            (our CUDA path)
    python bench.py --impl reference ...                     (CPU arm: the oracle port, --cpu-threads host threads)

A step = one pass of the hot path over one batch: reset -> fused trim+pileup kernel -> link+call kernels.
N=1 workload = BASELINE.json configs[1]: 1M Illumina 2x150 reads, ARTIC-v3-like scheme, 29,903 bp.
N>1: weak scaling, one such sample per GPU (plate-style sample sharding, no data-path collective).

  value  = whole-job reads/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e    = the same through the host-buffer C-ABI call (pinned host arrays; H2D/D2H inside the timed region)
  roofline.achieved = algorithmic bytes of the fused kernel / its average launch duration (events per launch)

Sub-objects next to the headline (same JSON line):
  ont      (N=1)  configs[3] shape: ONT-like ~410-base reads with ~47 CIGAR ops, device-resident step + kernel roofline
  file_e2e (N=1)  `aio` through the command line on files: BAM in -> trimmed BAM + VCF + FASTA out (warm, best of 3)
  deep     (N>1)  configs[2] shape: ONE deep sample (the N=1 sample at --deep-copies x depth), its reads sharded over the
                  ranks, then ncclAllReduce of the count matrix + packed insertion-table exchange + calling; strong scaling;
                  checked against the same sample processed on one GPU
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L_GENOME = 29903
METRIC = "aligned reads/sec trimmed+piled-up"
FIELDS = ("pos", "flag", "tlen", "cig_off", "cigar", "seq_off", "seq", "qual_off", "qual")


def make_workload(n_reads, seed, kind="illumina"):
    from amplipy_b200 import synth
    from amplipy_b200.batch import ReadBatch
    g = synth.random_genome(L_GENOME, 7)
    primers, amps = synth.make_scheme(L_GENOME, 98, seed=2)
    prim = [(s, e) for s, e, _ in primers]
    snvs = [(1000 + 2800 * i, "ACGT"[i % 4], af) for i, af in enumerate([1.0, 0.9, 0.75, 0.5, 0.4, 0.3, 0.2, 0.1, 0.05, 0.05])]
    cache = "/tmp/amplipy_b200_bench_%s_%d_%d.npz" % (kind, n_reads, seed)
    b = None
    if os.path.isfile(cache):
        try:
            z = np.load(cache)
            b = ReadBatch(*[z[f] for f in FIELDS])
        except Exception:
            b = None
    if b is None:
        if kind == "ont":
            b = synth.ont_batch(g, amps, n_reads, seed=seed)
        else:
            b = synth.illumina_batch(g, amps, n_reads, seed=seed, snvs=snvs)
        try:
            tmp = "%s.%d.tmp.npz" % (cache, os.getpid())
            np.savez(tmp, **{f: getattr(b, f) for f in FIELDS})
            os.replace(tmp, cache)
        except OSError:
            pass
    return g, prim, b


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (pynvml, 20 Hz)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def ncu_traffic(kind, n_reads):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed `ncu --set full` capture
    (profiles/traffic.json; per launch of exactly this workload), else None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = d.get(kind, d if kind == "illumina" else {})
        return float(e["traffic_bytes_per_launch"]) if int(e.get("reads", 1_000_000)) == n_reads else None
    except Exception:
        return None


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_config(args):
    """Identical in both arms (the driver compares the two lines' config objects)."""
    return {"workload": "configs[1]: synthetic SARS-CoV-2-length genome (29,903 bp, seeded random stand-in), ARTIC-v3-like "
                        "98-amplicon scheme (generated), %d %s reads per GPU, coordinate-sorted" %
                        (args.reads, "Illumina 2x150" if args.workload == "illumina" else "ONT ~400bp"),
            "reads_per_gpu": args.reads, "ref_len": L_GENOME, "min_quality": 20, "sliding_window": 4, "min_length": 30,
            "l2": "inputs larger than L2 (about 250 MB read per step vs 126 MB L2)",
            "sharding": "sample per GPU, no collective"}


def cpu_oracle(args):
    """The CPU arm's library: the oracle port built -O3 -march=native on this machine, on a fixed number of threads
    (whatever OMP_NUM_THREADS says -- torchrun sets it to 1)."""
    from oracle import oracle
    oracle.use_native_build()
    oracle.set_num_threads(args.cpu_threads)
    return oracle


def oracle_pass(oracle, b, g, prim, tables, mpl):
    """One CPU pass of the same work with the oracle port: trim -> pileup -> call."""
    t = oracle.trim_batch(b, L_GENOME, tables[0], tables[1], mpl)
    counts, ins, _ = oracle.pileup_batch(b, L_GENOME, 20, trimmed=t)
    res = oracle.call(counts, ins, g)
    return t, counts, ins, res


def run_reference(args, rank, world):
    """CPU arm: the reference is pure Python and cannot be compiled into oracle/_ref, so this times the
    oracle port (C + OpenMP, --cpu-threads host threads) on a bounded sample of the same workload."""
    if rank != 0:
        return
    oracle = cpu_oracle(args)
    from amplipy_b200.primers import find_overlapping_primers, max_primer_len
    sample = min(args.reads, args.cpu_sample)
    g, prim, b = make_workload(args.reads, args.seed, args.workload)
    b = b.slice(0, sample)
    tables = find_overlapping_primers(L_GENOME, prim, 0)
    mpl = max_primer_len(prim)
    for _ in range(args.warmup):
        oracle_pass(oracle, b, g, prim, tables, mpl)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_pass(oracle, b, g, prim, tables, mpl)
    dt = (time.perf_counter() - t0) / args.steps
    v = sample / dt
    cores = oracle.num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "port",
                             "sample": "first %d reads of the workload, oracle/amplipy_oracle.c (C+OpenMP restatement of "
                                       "AmpliPy.py, -O3 -march=native), %d threads" % (sample, cores),
                             "amplipy_py_note": "the unmodified AmpliPy.py (pure Python over the pysam shim) runs the same "
                                                "trim+pileup+call at 11.7 k reads/s on one core (measured in the build container, BASELINE.md); it has no "
                                                "multi-threaded mode"},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def pin_batch(b):
    """Page-locked copies of a batch's arrays (torch owns the memory; numpy views are handed to the C ABI)."""
    import torch
    from amplipy_b200.batch import ReadBatch
    keep, views = [], []
    for f in FIELDS:
        a = getattr(b, f)
        v = a.view(np.int16) if a.dtype == np.uint16 else a.view(np.int32) if a.dtype == np.uint32 else a
        t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
        keep.append(t)
        views.append(t.numpy().view(a.dtype))
    return ReadBatch(*views), keep


def pin_outputs(outs):
    import torch
    keep, views = [], []
    for a in outs:
        v = a.view(np.int16) if a.dtype == np.uint16 else a.view(np.int32) if a.dtype == np.uint32 else a
        t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
        keep.append(t)
        views.append(t.numpy().view(a.dtype))
    return tuple(views), keep


def live_rows(b, ncig, n):
    """indices of the live words of the first n trim-output CIGAR rows (row i at cig_off[i] + 3 i, ncig[i] words)"""
    nc = ncig[:n].astype(np.int64)
    row0 = b.cig_off[:n].astype(np.int64) + 3 * np.arange(n, dtype=np.int64)
    return np.repeat(row0, nc) + (np.arange(int(nc.sum())) - np.repeat(np.cumsum(nc) - nc, nc))


def parity_on_sample(eng, oracle, sb, g, prim, tables, mpl, pouts, sample):
    """The CPU baseline run doubles as a parity check of the benchmarked path on the sample: trim outputs (positions,
    flags, CIGAR rows) of the timed end-to-end step, and counts, insertion alleles, depth and consensus of the sample."""
    from amplipy_b200 import calling
    ot, oc, oi, ores = oracle_pass(oracle, sb, g, prim, tables, mpl)
    ok = {"pos": bool(np.array_equal(pouts[0][:sample], ot["pos"])),
          "flags": bool(np.array_equal(pouts[2][:sample], ot["flags"])),
          "ncig": bool(np.array_equal(pouts[1][:sample].astype(np.int32), ot["ncig"]))}
    live = live_rows(sb, ot["ncig"], sample)
    ok["cigar"] = bool(np.array_equal(pouts[3][live], ot["cigar"][live]))
    eng.reset()
    eng.process(sb, trim=True, pileup=True)
    counts, ins = eng.counts(), eng.insertions()
    ok["counts"] = bool(np.array_equal(counts.astype(np.int64), oc))
    ok["insertions"] = bool(ins.as_dict() == oi)
    res = eng.call(g)
    ok["depth"] = bool(np.array_equal(res.depth.astype(np.int64), ores["depth"]))
    ok["consensus"] = bool(calling.consensus_string(res, ins) == oracle.consensus_string(ores))
    return all(ok.values()), ok


def file_to_file(args, g, prim, b, runs=5):
    """`aio` through the command line on files: BAM bytes -> GPU (inflate, trim + pileup, calling, record rebuild, deflate) -> trimmed
    BAM + VCF + FASTA.  Reported next to the kernel numbers because this is what a real run sees.  Warm: the CUDA context and the
    libraries are up; best of `runs`."""
    import shutil
    import tempfile
    from amplipy_b200 import alnio, cli, synth
    d = tempfile.mkdtemp(prefix="amplipy_b200_bench_")
    try:
        j = lambda n: os.path.join(d, n)
        hdr = "@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:ref\tLN:%d\n@PG\tID:synth\tPN:synth\n" % len(g)
        alnio.write_bam(j("in.bam"), hdr, [("ref", len(g))], b)
        synth.write_bed(j("primers.bed"), [(s, e, "p%d" % k) for k, (s, e) in enumerate(prim)])
        synth.write_fasta(j("ref.fas"), "ref", g)
        t0 = time.perf_counter()
        a = alnio.read_alignments(j("in.bam"))
        t_decode = time.perf_counter() - t0
        assert a.batch.n == b.n
        del a
        def one_run():
            for f in ("trimmed.bam", "variants.vcf", "consensus.fas", "timings.json"):
                if os.path.exists(j(f)):
                    os.remove(j(f))
            os.environ["AMP_CLI_TIMINGS"] = j("timings.json")
            t0 = time.perf_counter()
            cli.main(["aio", "-i", j("in.bam"), "-p", j("primers.bed"), "-r", j("ref.fas"), "-ot", j("trimmed.bam"),
                      "-ov", j("variants.vcf"), "-oc", j("consensus.fas")])
            dt = time.perf_counter() - t0
            try:
                sp = json.load(open(j("timings.json")))
            except Exception:
                sp = None
            return dt, sp
        times, splits = [], []
        for r in range(runs):
            dt, sp = one_run()
            times.append(dt); splits.append(sp)
        size_dev = os.path.getsize(j("trimmed.bam"))
        os.environ["AMPLIPY_BAM_LEVEL"] = "6"           # the trimmed BAM rebuilt and deflated on the host with zlib at htslib's default level
        slow = [one_run() for _ in range(2)]
        os.environ.pop("AMPLIPY_BAM_LEVEL", None)
        kf = int(np.argmin([x[0] for x in slow]))
        host_obj = {"value": b.n / slow[kf][0], "unit": "reads/s", "seconds": slow[kf][0], "split_s": slow[kf][1],
                    "trimmed_bam_bytes": os.path.getsize(j("trimmed.bam")), "note": "AMPLIPY_BAM_LEVEL=6: host inflate + record rewrite + zlib"}
        os.environ.pop("AMP_CLI_TIMINGS", None)
        k = int(np.argmin(times))
        return {"value": b.n / times[k], "unit": "reads/s", "seconds": times[k], "runs_s": times, "split_s": splits[k],
                "bam_decode_only_reads_per_s": b.n / t_decode,
                "bam_bytes": os.path.getsize(j("in.bam")), "trimmed_bam_bytes": size_dev, "host_zlib_level_6": host_obj,
                "note": "python -m amplipy_b200 aio on files in a temp directory (page cache warm), best of %d: the BAM is inflated, "
                        "trimmed, rebuilt and deflated on the GPU (the device compressor writes files a little smaller than zlib level 1); "
                        "the reference's counterpart is AmpliPy.py aio with pysam I/O" % runs}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def ont_leg(args, local_rank):
    """configs[3] shape on one GPU: ONT-like reads (about 410 bases, 47 CIGAR ops), every read on the general CIGAR path."""
    import torch
    from amplipy_b200.engine import Engine
    from amplipy_b200.primers import find_overlapping_primers, max_primer_len
    g, prim, b = make_workload(args.ont_reads, 4, "ont")
    tables = find_overlapping_primers(L_GENOME, prim, 0)
    eng = Engine(ref_len=L_GENOME, primer_tables=tables, max_primer_len=max_primer_len(prim), device=local_rank,
                 ins_slots=1 << 24, ins_arena_bytes=1 << 30)
    eng.set_reference(g)
    stream = torch.cuda.Stream()
    d = eng.upload(b)
    eng.reserve(b.n, int(b.cig_off[-1]))
    steps = max(3, min(args.steps, 10))
    with torch.cuda.stream(stream):
        for _ in range(3):
            eng.reset_async(stream.cuda_stream)
            eng.process_device(d, stream=stream.cuda_stream)
            eng.call_device(stream=stream.cuda_stream)
        torch.cuda.synchronize()
        out_bytes = b.n * 7 + 4 * int(d["o_ncig"].to(torch.int64).sum().item())
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ka = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        kb = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ev0.record(stream)
        for k in range(steps):
            eng.reset_async(stream.cuda_stream)
            ka[k].record(stream)
            eng.process_device(d, stream=stream.cuda_stream)
            kb[k].record(stream)
            eng.call_device(stream=stream.cuda_stream)
        ev1.record(stream)
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    kern_ms = float(np.mean([a.elapsed_time(c) for a, c in zip(ka, kb)]))
    peak, _ = peak_hbm()
    alg = b.algorithmic_bytes() + out_bytes
    flags = eng.error_flags()
    n_ins = int(eng.insertions().k) if hasattr(eng.insertions(), "k") else None
    eng.close()
    return {"workload": "configs[3] shape: %d ONT-like single-end reads (about 410 bases, %.0f CIGAR ops per read, 3%% ins / 3%% del), "
                        "coordinate-sorted" % (b.n, float(b.cig_off[-1]) / b.n),
            "reads": b.n, "ms_per_step": ms, "reads_per_s": b.n / (ms / 1e3), "kernel_ms": kern_ms,
            "bytes_per_read": alg / b.n, "achieved_gbs": alg / (kern_ms / 1e3) / 1e9, "frac": alg / (kern_ms / 1e3) / 1e9 / peak,
            "traffic": ncu_traffic("ont", b.n), "distinct_insertion_alleles": n_ins, "device_error_flags": flags}


def deep_leg(args, rank, world, local_rank, stream):
    """configs[2] shape, strong scaling: ONE deep sample = the N=1 sample with every read --deep-copies times (coordinate
    order kept), read ranges per rank, then the exchange step (ncclAllReduce of the count matrix, packed insertion-table
    all-gather + merge) and calling on every rank.  Checked against the base sample processed alone on this GPU: counts,
    insertion-allele counts and depth must be exactly deep-copies times the base's, and every rank must hold the same totals."""
    import torch
    import torch.distributed as dist
    from amplipy_b200 import dist as adist
    from amplipy_b200 import synth
    from amplipy_b200.engine import Engine
    from amplipy_b200.primers import find_overlapping_primers, max_primer_len
    copies = args.deep_copies
    g, prim, base = make_workload(args.reads, args.seed, "illumina")
    tables = find_overlapping_primers(L_GENOME, prim, 0)
    mpl = max_primer_len(prim)
    first, count = adist.read_range(base.n, rank, world)
    sl = base.slice(first, first + count)
    mine = synth._reorder(sl, np.repeat(np.arange(sl.n, dtype=np.int64), copies))
    eng = Engine(ref_len=L_GENOME, primer_tables=tables, max_primer_len=mpl, device=local_rank)
    eng.set_reference(g)
    d = eng.upload(mine)
    eng.reserve(mine.n, int(mine.cig_off[-1]))
    ex = adist.DeepExchange(eng, cap_entries=1 << 15, cap_arena_bytes=2 << 20)
    s = stream.cuda_stream
    steps = max(3, min(args.steps, 20))

    def one(evs=None):
        eng.reset_async(s)
        if evs: evs[0].record(stream)
        eng.process_device(d, stream=s)
        if evs: evs[1].record(stream)
        eng.allreduce_counts(ex.comm, s)
        if evs: evs[2].record(stream)
        eng.ins_pack_device(ex.send.data_ptr(), ex.cap_entries, ex.cap_arena, s)
        ex.comm.allgather(ex.send.data_ptr(), ex.recv.data_ptr(), ex.slot_bytes, s)
        eng.ins_merge_packed(ex.recv.data_ptr(), world, rank, ex.cap_entries, ex.cap_arena, s)
        if evs: evs[3].record(stream)
        eng.call_device(stream=s)
        if evs: evs[4].record(stream)

    with torch.cuda.stream(stream):
        for _ in range(3):
            one()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(steps):
            one(evs[k])
        e1.record(stream)
        dist.barrier()
        torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1) / steps
    parts = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(4)] for e in evs]).mean(axis=0)
    tt = torch.tensor([total_ms] + parts.tolist(), device="cuda", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, kern_ms, ar_ms, ins_ms, call_ms = [float(x) for x in tt.tolist()]
    # ---- parity: every rank against the base sample processed alone on its own GPU
    counts = eng.counts().astype(np.int64)
    ins = eng.insertions().as_dict()
    res = eng.call(g)
    flags = eng.error_flags()
    solo = Engine(ref_len=L_GENOME, primer_tables=tables, max_primer_len=mpl, device=local_rank)
    solo.process(base, trim=True, pileup=True)
    bc = solo.counts().astype(np.int64)
    bi = solo.insertions().as_dict()
    bres = solo.call(g)
    ok = bool(np.array_equal(counts, bc * copies) and ins == {k: v * copies for k, v in bi.items()} and
              np.array_equal(res.depth.astype(np.int64), bres.depth.astype(np.int64) * copies) and flags == 0)
    h = hashlib.sha256(counts.tobytes() + repr(sorted(ins.items())).encode()).digest()[:8]
    chk = torch.tensor([int(counts.sum()), len(ins), int.from_bytes(h, "little") >> 1, int(ok)], device="cuda", dtype=torch.int64)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(lo, hi))
    ex.close(); eng.close(); solo.close()
    n_total = base.n * copies
    return {"workload": "configs[2] shape: one deep sample, %d reads (the N=1 sample at %dx depth, about %d,000x coverage), "
                        "contiguous read ranges per rank" % (n_total, copies, int(round(n_total * 125 / L_GENOME / 1000.0))),
            "reads": n_total, "ranks": world, "scaling": "strong", "ms_per_step": total_ms, "reads_per_s": n_total / (total_ms / 1e3),
            "kernel_ms": kern_ms, "allreduce_ms": ar_ms, "ins_exchange_ms": ins_ms, "call_ms": call_ms,
            "exchange": "amp_allreduce_counts (ncclAllReduce, int32 sum, %d bytes) + amp_ins_pack_device -> ncclAllGather of %d-byte slots "
                        "-> amp_ins_merge_packed; no host round trip" % (6 * eng.lpad * 4, ex.slot_bytes),
            "deep_parity": bool(ok and same and int(lo[3].item()) == 1),
            "checksum": {"sum_counts": int(counts.sum()), "insertion_alleles": len(ins), "ranks_agree": same}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--workload", default="illumina", choices=["illumina", "ont"])
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=200_000)
    ap.add_argument("--cpu-threads", type=int, default=os.cpu_count() or 1,
                    help="threads of the CPU arm / cpu_baseline leg (set explicitly: torchrun exports OMP_NUM_THREADS=1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps (capped at 10)")
    ap.add_argument("--no-file-e2e", action="store_true", help="skip the file-to-file command line leg (N=1)")
    ap.add_argument("--file-e2e", action="store_true", help="(kept for compatibility: the leg runs by default at N=1)")
    ap.add_argument("--no-ont", action="store_true", help="skip the ONT-like sub-object (N=1)")
    ap.add_argument("--ont-reads", type=int, default=300_000)
    ap.add_argument("--no-deep", action="store_true", help="skip the deep-sample sub-object (N>1)")
    ap.add_argument("--deep-copies", type=int, default=8)
    ap.add_argument("--lean", action="store_true", help="headline numbers only (tuning runs): no cpu baseline, ont, file or deep legs")
    args = ap.parse_args()
    if args.lean:
        args.no_cpu_baseline = args.no_file_e2e = args.no_ont = args.no_deep = True
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]           # keep NCCL's version banner off stdout: rank 0 prints one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from amplipy_b200.engine import Engine
    from amplipy_b200.primers import find_overlapping_primers, max_primer_len

    g, prim, b = make_workload(args.reads, args.seed + rank, args.workload)
    tables = find_overlapping_primers(L_GENOME, prim, 0)
    mpl = max_primer_len(prim)
    ont = args.workload == "ont"
    eng = Engine(ref_len=L_GENOME, primer_tables=tables, max_primer_len=mpl, device=local_rank,
                 ins_slots=(1 << 24) if ont else 0, ins_arena_bytes=(1 << 30) if ont else 0)
    eng.set_reference(g)
    stream = torch.cuda.Stream()
    d = eng.upload(b)
    eng.reserve(b.n, int(b.cig_off[-1]))
    in_bytes = b.algorithmic_bytes()

    bm = os.environ.get("AMP_BENCH_MODE", "aio")     # tuning experiments only: time one half of the fused kernel
    do_trim, do_pile = bm in ("aio", "trim"), bm in ("aio", "pileup")

    def step_device():
        eng.reset_async(stream.cuda_stream)
        eng.process_device(d, trim=do_trim, pileup=do_pile, stream=stream.cuda_stream)
        eng.call_device(stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_device()
        torch.cuda.synchronize()
        # algorithmic output bytes of the fused kernel (exact, from the warm-up's own outputs)
        out_ncig = int(d["o_ncig"].to(torch.int64).sum().item())
        out_bytes = b.n * (4 + 2 + 1) + 4 * out_ncig
        eng.launches = 0
        sampler = ClockSampler(local_rank)
        sampler.start()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ka = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        kb = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        ev0.record(stream)
        for k in range(args.steps):
            eng.reset_async(stream.cuda_stream)
            ka[k].record(stream)
            eng.process_device(d, trim=do_trim, pileup=do_pile, stream=stream.cuda_stream)
            kb[k].record(stream)
            eng.call_device(stream=stream.cuda_stream)
        ev1.record(stream)
        barrier()
        sampler.stop_flag = True
        total_ms = ev0.elapsed_time(ev1)
        kern_ms = float(np.mean([a.elapsed_time(c) for a, c in zip(ka, kb)]))
        launches = eng.launches
    if world > 1:
        tt = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps
    value = b.n * world / (ms_per_step / 1e3)
    flags_dev = eng.error_flags()

    # ---- e2e: host buffers through amp_process_host + amp_call --------------------------------------
    pb, keep_in = pin_batch(b)
    pouts, keep_out = pin_outputs(Engine.alloc_trim_out(pb))
    e2e_steps = args.e2e_steps or min(args.steps, 10)

    def step_e2e():
        eng.reset()
        eng.process(pb, trim=True, pileup=True, out=pouts)
        return eng.call(None, pinned=True)

    for _ in range(2):
        res = step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = step_e2e()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        tt = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_soa = {"value": b.n * world / e2e_s, "unit": "reads/s", "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "h2d_bytes_per_step": eng.host_copy_bytes(b)[0], "d2h_bytes_per_step": eng.host_copy_bytes(b)[1],
               "note": "amp_process_host: decoded struct-of-arrays batch in pinned host memory -> chunked H2D -> fused kernel -> D2H "
                       "trim outputs, + amp_call"}

    # ---- e2e, headline: the BAM file's bytes through amp_bam_decode_host + amp_process_decoded + amp_call ----------------
    # (what the reference's pysam layer reads; the compressed file crosses PCIe as it is and is decoded in HBM)
    import tempfile
    from amplipy_b200 import alnio
    with tempfile.TemporaryDirectory(prefix="amplipy_b200_bench_") as td:
        bam_path = os.path.join(td, "in.bam")
        alnio.write_bam(bam_path, "@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:ref\tLN:%d\n@PG\tID:synth\tPN:synth\n" % len(g), [("ref", len(g))], b,
                        level=6)
        bam_raw = np.fromfile(bam_path, np.uint8)
    layout = alnio.bam_layout(bam_raw.tobytes())
    bam_pin_t = torch.from_numpy(bam_raw).pin_memory()
    bam_pin = bam_pin_t.numpy()

    def step_bam():
        eng.reset()
        eng.decode_bam(bam_pin, layout)
        eng.process_decoded(trim=True, pileup=True, out=pouts)
        return eng.call(None, pinned=True)

    for _ in range(2):
        res = step_bam()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = step_bam()
    torch.cuda.synchronize()
    bam_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        tt = torch.tensor([bam_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        bam_s = float(tt.item())
    e2e_bam = {"value": b.n * world / bam_s, "unit": "reads/s", "ms_per_step": bam_s * 1e3, "steps": e2e_steps,
               "h2d_bytes_per_step": int(bam_raw.size) + 20 * int(layout["in_off"].size), "d2h_bytes_per_step": eng.host_copy_bytes(b)[1],
               "note": "the BAM file's bytes (pinned host memory, %d bytes for %d reads) -> amp_bam_decode_host (H2D as is, inflate + "
                       "record scatter on the device) -> amp_process_decoded (fused kernel, D2H trim outputs) -> amp_call (D2H call "
                       "outputs)" % (int(bam_raw.size), b.n)}
    # the headline e2e is the faster of the two host-buffer entry points
    e2e_pick = e2e_bam if e2e_bam["value"] > e2e_soa["value"] else e2e_soa

    peak, peak_src = peak_hbm()
    achieved = (in_bytes + out_bytes) / (kern_ms / 1e3) / 1e9
    kname = "amp_trim_pileup_warp_kernel" if not ont else "amp_trim_pileup_indel_kernel"
    line = {"metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": workload_config(args),
            "e2e": e2e_pick, "e2e_soa": e2e_soa, "e2e_bam": e2e_bam,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(args.workload, b.n), "kernel": kname, "kernel_ms": kern_ms,
                         "algorithmic_bytes_per_launch": in_bytes + out_bytes, "bytes_per_read": (in_bytes + out_bytes) / b.n,
                         "peak_source": peak_src, "kernel_share_of_step": kern_ms / ms_per_step},
            "clocks": sampler.summary(), "device_error_flags": flags_dev,
            "depth_checksum": int(res.depth.astype(np.int64).sum())}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        oracle = cpu_oracle(args)
        sample = min(b.n, args.cpu_sample)
        sb = b.slice(0, sample)
        oracle_pass(oracle, sb, g, prim, tables, mpl)
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            oracle_pass(oracle, sb, g, prim, tables, mpl)
        dt = (time.perf_counter() - t0) / reps
        cores = oracle.num_threads()
        same, detail = parity_on_sample(eng, oracle, sb, g, prim, tables, mpl, pouts, sample)
        line["cpu_baseline"] = {"value": sample / dt, "unit": "reads/s", "cores": cores, "kind": "port",
                                "sample": "first %d reads of the workload; oracle/amplipy_oracle.c (C+OpenMP restatement of "
                                          "AmpliPy.py trim+pileup+call, -O3 -march=native), %d threads" % (sample, cores),
                                "parity_on_sample": same, "parity_detail": detail}
    if world == 1 and not args.no_ont and not ont:
        line["ont"] = ont_leg(args, local_rank)
    if rank == 0 and world == 1 and not args.no_file_e2e:
        line["file_e2e"] = file_to_file(args, g, prim, b)
    if world > 1 and not args.no_deep and not ont:
        line["deep"] = deep_leg(args, rank, world, local_rank, stream)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
