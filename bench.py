#!/usr/bin/env python3
"""bench.py -- reads/s trimmed + piled-up (+ called) on synthetic SARS-CoV-2-shaped amplicon data.

    python bench.py --gpus N --steps K --warmup W            (our CUDA path)
    python bench.py --impl reference ...                     (CPU arm: the oracle port, all host threads)

A step = one pass of the hot path over one batch: reset -> fused trim+pileup kernel -> link+call kernels.
N=1 workload = BASELINE.json configs[1]: 1M Illumina 2x150 reads, ARTIC-v3-like scheme, 29,903 bp.
N>1: weak scaling, one such sample per GPU (plate-style sample sharding, no data-path collective).

  value  = whole-job reads/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e    = the same through the host-buffer C-ABI call (pinned host arrays; H2D/D2H inside the timed region)
  roofline.achieved = algorithmic bytes of the fused kernel / its average launch duration (events per launch)
"""
import argparse
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


def make_workload(n_reads, seed, kind="illumina"):
    from amplipy_b200 import synth
    g = synth.random_genome(L_GENOME, 7)
    primers, amps = synth.make_scheme(L_GENOME, 98, seed=2)
    prim = [(s, e) for s, e, _ in primers]
    snvs = [(1000 + 2800 * i, "ACGT"[i % 4], af) for i, af in enumerate([1.0, 0.9, 0.75, 0.5, 0.4, 0.3, 0.2, 0.1, 0.05, 0.05])]
    cache = "/tmp/amplipy_b200_bench_%s_%d_%d.npz" % (kind, n_reads, seed)
    from amplipy_b200.batch import ReadBatch
    fields = ("pos", "flag", "tlen", "cig_off", "cigar", "seq_off", "seq", "qual_off", "qual")
    if os.path.isfile(cache):
        z = np.load(cache)
        b = ReadBatch(*[z[f] for f in fields])
    else:
        if kind == "ont":
            b = synth.ont_batch(g, amps, n_reads, seed=seed)
        else:
            b = synth.illumina_batch(g, amps, n_reads, seed=seed, snvs=snvs)
        try:
            np.savez(cache, **{f: getattr(b, f) for f in fields})
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


def ncu_traffic(n_reads):
    """dram__bytes_read.sum + dram__bytes_write.sum of the fused kernel from the committed `ncu --set full` capture
    (profiles/traffic.json, taken on the 1M-read config-2 launch); None for other workload sizes."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(p))
        return float(d["traffic_bytes_per_launch"]) if n_reads == 1_000_000 else None
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


def oracle_pass(oracle, b, g, prim, tables, mpl):
    """One CPU pass of the same work with the oracle port: trim -> pileup -> call."""
    t = oracle.trim_batch(b, L_GENOME, tables[0], tables[1], mpl)
    counts, ins, _ = oracle.pileup_batch(b, L_GENOME, 20, trimmed=t)
    oracle.call(counts, ins, g)
    return t, counts, ins


def run_reference(args, rank, world):
    """CPU arm: the reference is pure Python and cannot be compiled into oracle/_ref, so this times the
    oracle port (C + OpenMP, every host thread) on a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import oracle
    oracle.build()
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
            "config": workload_config(args, args.reads),
            "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "port",
                             "sample": "first %d reads of the workload, oracle/amplipy_oracle.c (C+OpenMP restatement of AmpliPy.py), %d threads" % (sample, cores)},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def file_to_file(args, g, prim, b):
    """`aio` through the command line on files: BAM decode (host, multi-threaded BGZF) -> GPU path -> BAM encode + VCF +
    FASTA.  Reported next to the kernel numbers because this is where a real run is bounded (host I/O)."""
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
        t0 = time.perf_counter()
        cli.main(["aio", "-i", j("in.bam"), "-p", j("primers.bed"), "-r", j("ref.fas"), "-ot", j("trimmed.bam"),
                  "-ov", j("variants.vcf"), "-oc", j("consensus.fas")])
        t_all = time.perf_counter() - t0
        return {"value": b.n / t_all, "unit": "reads/s", "seconds": t_all, "bam_decode_only_reads_per_s": b.n / t_decode,
                "bam_bytes": os.path.getsize(j("in.bam")), "trimmed_bam_bytes": os.path.getsize(j("trimmed.bam")),
                "note": "python -m amplipy_b200 aio on files in a temp directory; one run, includes process-level set-up of the engine"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def workload_config(args, n_reads):
    return {"workload": "configs[1]: synthetic SARS-CoV-2-length genome (29,903 bp, seeded random stand-in), ARTIC-v3-like "
                        "98-amplicon scheme (generated), %d %s reads per GPU, coordinate-sorted" %
                        (n_reads, "Illumina 2x150" if args.workload == "illumina" else "ONT ~400bp"),
            "reads_per_gpu": n_reads, "ref_len": L_GENOME, "min_quality": 20, "sliding_window": 4, "min_length": 30,
            "l2": "inputs larger than L2 (about 250 MB read per step vs 126 MB L2)", "sharding": "sample per GPU, no collective"}


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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps (capped at 10)")
    ap.add_argument("--file-e2e", action="store_true",
                    help="also time the command line file to file (BAM in -> trimmed BAM + VCF + FASTA out) and add a file_e2e object")
    args = ap.parse_args()
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
    from amplipy_b200.batch import ReadBatch
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
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    keep = []
    hb = []
    for f in ("pos", "flag", "tlen", "cig_off", "cigar", "seq_off", "seq", "qual_off", "qual"):
        a = getattr(b, f)
        if a.dtype == np.uint16:
            t, v = pin(a.view(np.int16)); v = v.view(np.uint16)
        elif a.dtype == np.uint32:
            t, v = pin(a.view(np.int32)); v = v.view(np.uint32)
        else:
            t, v = pin(a)
        keep.append(t); hb.append(v)
    pb = ReadBatch(*hb)
    outs = Engine.alloc_trim_out(pb)
    pouts = []
    for a in outs:
        if a.dtype == np.uint16:
            t, v = pin(a.view(np.int16)); v = v.view(np.uint16)
        elif a.dtype == np.uint32:
            t, v = pin(a.view(np.int32)); v = v.view(np.uint32)
        else:
            t, v = pin(a)
        keep.append(t); pouts.append(v)
    pouts = tuple(pouts)
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
    e2e_value = b.n * world / e2e_s
    d2h = out_bytes + 4 * 3 * b.n * 0 + L_GENOME * (4 + 4 + 4 + 1 + 4 + 48 + 24 + 1)
    d2h = b.n * (4 + 2 + 1) + 4 * (int(b.cig_off[-1]) + 3 * b.n) + L_GENOME * (4 + 4 + 4 + 1 + 4 + 48 + 24 + 1)

    peak, peak_src = peak_hbm()
    achieved = (in_bytes + out_bytes) / (kern_ms / 1e3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": workload_config(args, b.n),
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                    "note": "amp_process_host (pinned host SoA -> chunked H2D -> fused kernel -> D2H trim outputs) + amp_call (D2H call outputs)"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(b.n) if args.workload == "illumina" else None, "kernel": "amp_trim_pileup_warp_kernel" if args.workload == "illumina" else "amp_trim_pileup_indel_kernel", "kernel_ms": kern_ms,
                         "algorithmic_bytes_per_launch": in_bytes + out_bytes, "bytes_per_read": (in_bytes + out_bytes) / b.n,
                         "peak_source": peak_src, "kernel_share_of_step": kern_ms / ms_per_step},
            "clocks": sampler.summary(), "device_error_flags": flags_dev,
            "depth_checksum": int(res.depth.astype(np.int64).sum())}
    line["config"]["l2"] = "inputs larger than L2 (%.0f MB read per step vs 126 MB L2)" % (in_bytes / 1e6)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        oracle.build()
        sample = min(b.n, args.cpu_sample)
        sb = b.slice(0, sample)
        oracle_pass(oracle, sb, g, prim, tables, mpl)
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            ot, oc, oi = oracle_pass(oracle, sb, g, prim, tables, mpl)
        dt = (time.perf_counter() - t0) / reps
        cores = oracle.num_threads()
        # the baseline run doubles as a parity check of the benchmarked outputs on the sample
        same = bool(np.array_equal(pouts[0][:sample], ot["pos"]) and np.array_equal(pouts[2][:sample], ot["flags"]))
        line["cpu_baseline"] = {"value": sample / dt, "unit": "reads/s", "cores": cores, "kind": "port",
                                "sample": "first %d reads of the workload; oracle/amplipy_oracle.c (C+OpenMP restatement of "
                                          "AmpliPy.py trim+pileup+call), %d threads" % (sample, cores),
                                "parity_on_sample": same}
    if rank == 0 and world == 1 and args.file_e2e:
        line["file_e2e"] = file_to_file(args, g, prim, b)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
