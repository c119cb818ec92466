"""Multi-GPU sharding of the path (one process per GPU, torch.distributed for the plumbing).

Two natural partitions:

* plate  -- samples are independent: sample i goes to rank i % world; no data-path collective at all.
* deep   -- one sample, contiguous read ranges per rank.  Trim outputs are per read (no exchange); each
            rank accumulates a private count matrix and insertion table; then ONE all-reduce (sum, int32)
            of the [6, lpad] matrix (NCCL over NVLink/NVSwitch on GPUs) and an all-gather of the distinct
            insertion alleles, which every rank merges into its own table with the merge kernel.
            Integer sums => the result is independent of the partition and bit-identical to one GPU.

The reference has no counterpart (single sequential loop, AmpliPy.py:896)."""
import numpy as np


def plate_assignment(n_samples, rank, world):
    """Samples owned by `rank` (round-robin)."""
    return list(range(rank, n_samples, world))


def read_range(n_reads, rank, world):
    """Contiguous [first, first + count) slice of a coordinate-sorted batch for `rank`."""
    first = (n_reads * rank) // world
    last = (n_reads * (rank + 1)) // world
    return first, last - first


def allreduce_counts(engine, group=None):
    """Sum the per-rank count matrices in place (every rank ends up with the total)."""
    import torch.distributed as dist
    t = engine.counts_tensor()
    if t.is_cuda:
        import torch
        torch.cuda.synchronize(t.device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allgather_insertions(engine, group=None):
    """Exchange the distinct insertion alleles of every rank and merge the others' into the local table."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    ins = engine.insertions()
    mine = (np.asarray(ins.sample), np.asarray(ins.pos), np.asarray(ins.count), np.asarray(ins.str_off), np.asarray(ins.chars))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    for r, (sample, pos, count, off, chars) in enumerate(gathered):
        if r != rank and len(pos):
            engine.merge_insertions(sample, pos, count, off, chars)
    return engine.insertions()


def process_deep_sample(engine, batch, trim=True, pileup=True, sample=0, group=None):
    """Read-sharded processing of one sample: local kernel on this rank's read range, then the exchange step.
    Returns this rank's TrimResult (valid for its own range of reads only)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    first, count = read_range(batch.n, rank, world)
    t = engine.process(batch, trim=trim, pileup=pileup, sample=sample, first=first, n=count)
    if pileup and world > 1:
        allreduce_counts(engine, group)
        allgather_insertions(engine, group)
    return t, (first, count)
