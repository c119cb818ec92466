"""Run under torchrun on >= 2 GPUs (NCCL): plate sharding and read-range sharding with the count all-reduce
must equal the single-GPU result (checked against the oracle on rank 0).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dist_gpu_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from amplipy_b200 import calling, synth
    from amplipy_b200 import dist as adist
    from amplipy_b200.engine import Engine
    from amplipy_b200.primers import find_overlapping_primers, max_primer_len
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = 29903
    g = synth.random_genome(L, 7)
    primers, amps = synth.make_scheme(L, 98, seed=2)
    prim = [(s, e) for s, e, _ in primers]
    tables = find_overlapping_primers(L, prim, 0)
    mpl = max_primer_len(prim)
    # ---- deep: one sample, read ranges per rank + all-reduce + insertion merge ------------------------------
    b = synth.illumina_batch(g, amps, 200_000, seed=77, p_ins=0.05)
    eng = Engine(ref_len=L, primer_tables=tables, max_primer_len=mpl, device=local)
    t, (first, count) = adist.process_deep_sample(eng, b)
    counts = eng.counts()
    ins = eng.insertions()
    res = eng.call(g)
    cons = calling.consensus_string(res, ins)
    if rank == 0:
        from oracle import oracle
        mn, mx = oracle.find_overlapping_primers(L, prim, 0)
        want = oracle.trim_batch(b, L, mn, mx, mpl)
        wc, wins, _ = oracle.pileup_batch(b, L, 20, trimmed=want)
        assert np.array_equal(counts.astype(np.int64), wc), "deep: counts differ from the oracle"
        assert ins.as_dict() == wins, "deep: insertion alleles differ from the oracle"
        assert cons == oracle.consensus_string(oracle.call(wc, wins, g))
        sl = slice(first, first + count)
        assert np.array_equal(t.pos[sl], want["pos"][sl]) and np.array_equal(t.flags[sl], want["flags"][sl])
    # every rank holds identical totals
    chk = torch.tensor([int(counts.astype(np.int64).sum()), len(ins.strs)], device="cuda", dtype=torch.int64)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "ranks disagree after the exchange"
    # ---- plate: independent samples, no collective ---------------------------------------------------------------
    n_samples = 2 * world
    mine = adist.plate_assignment(n_samples, rank, world)
    plate = Engine(ref_len=L, primer_tables=tables, max_primer_len=mpl, device=local, n_samples=len(mine))
    sums = []
    for k, s in enumerate(mine):
        sb = synth.illumina_batch(g, amps, 50_000, seed=1000 + s)
        plate.process(sb, sample=k)
        sums.append((s, int(plate.counts(k).astype(np.int64).sum())))
    allsums = [None] * world
    dist.all_gather_object(allsums, sums)
    if rank == 0:
        got = dict(x for part in allsums for x in part)
        assert sorted(got) == list(range(n_samples))
        solo = Engine(ref_len=L, primer_tables=tables, max_primer_len=mpl, device=local)
        for s in (0, n_samples - 1):
            solo.reset()
            solo.process(synth.illumina_batch(g, amps, 50_000, seed=1000 + s))
            assert int(solo.counts().astype(np.int64).sum()) == got[s]
        print("dist_gpu_check ok: world=%d deep+plate" % world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
