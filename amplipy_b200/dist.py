"""Multi-GPU sharding of the path (one process per GPU, torch.distributed for the plumbing).

Two natural partitions:

* plate  -- samples are independent: sample i goes to rank i % world; no data-path collective at all.
* deep   -- one sample, contiguous read ranges per rank.  Trim outputs are per read (no exchange); each
            rank accumulates a private count matrix and insertion table; then ONE all-reduce (sum, int32)
            of the [6, lpad] matrix over NVLink/NVSwitch and ONE all-gather of the ranks' insertion tables, packed on
            the device into fixed-size slots, merged into every rank's table by one kernel.  Nothing passes through
            the host.  Integer sums => the result is independent of the partition and bit-identical to one GPU.

On CUDA engines the exchange runs through the C ABI (amp_allreduce_counts, amp_ins_pack_device, amp_nccl_allgather,
amp_ins_merge_packed) on a communicator made by ``DeepExchange``; engines without that entry point (the CPU emulation
used by the gloo tests) take the torch.distributed form of the same two steps.

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


class DeepExchange:
    """Device-side exchange step of the deep-sample mode for one CUDA engine: an NCCL communicator of its own (created
    through the C ABI; the id travels over the torch.distributed group once) and the two slot buffers, allocated once.

    cap_entries / cap_arena_bytes: capacity of a rank's slot (distinct insertion alleles / bytes of their records); a
    table that does not fit raises the engine's TABLE_FULL / ARENA_FULL error flag."""

    def __init__(self, engine, group=None, cap_entries=1 << 16, cap_arena_bytes=4 << 20):
        import torch
        import torch.distributed as dist
        from .engine import NcclComm
        self.engine = engine
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.cap_entries, self.cap_arena = int(cap_entries), int(cap_arena_bytes)
        self.slot_bytes = engine.ins_slot_bytes(self.cap_entries, self.cap_arena)
        dev = torch.device("cuda", engine.device)
        self.send = torch.zeros(self.slot_bytes, dtype=torch.uint8, device=dev)
        self.recv = torch.zeros(self.slot_bytes * self.world, dtype=torch.uint8, device=dev)
        self.comm = NcclComm.from_torch_group(engine.device, group)

    def run(self, stream=None):
        """all-reduce of the counts + exchange of the insertion tables, asynchronous on ``stream`` (a raw cudaStream_t)."""
        e = self.engine
        e.allreduce_counts(self.comm, stream)
        e.ins_pack_device(self.send.data_ptr(), self.cap_entries, self.cap_arena, stream)
        self.comm.allgather(self.send.data_ptr(), self.recv.data_ptr(), self.slot_bytes, stream)
        e.ins_merge_packed(self.recv.data_ptr(), self.world, self.rank, self.cap_entries, self.cap_arena, stream)

    def close(self):
        self.comm.close()


def allreduce_counts(engine, group=None):
    """torch.distributed form: sum the per-rank count matrices in place (every rank ends up with the total)."""
    import torch.distributed as dist
    t = engine.counts_tensor()
    if t.is_cuda:
        import torch
        torch.cuda.synchronize(t.device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allgather_insertions(engine, group=None):
    """torch.distributed form (host arrays): exchange the distinct insertion alleles of every rank and merge the others'
    into the local table.  Used by engines without the packed device-side exchange (the CPU emulation)."""
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


def process_deep_sample(engine, batch, trim=True, pileup=True, sample=0, group=None, exchange=None):
    """Read-sharded processing of one sample: local kernel on this rank's read range, then the exchange step.
    Returns this rank's TrimResult (valid for its own range of reads only; the other reads' flags are 0).
    ``exchange``: a DeepExchange to reuse (CUDA engines); one is created and closed when it is None."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    first, count = read_range(batch.n, rank, world)
    t = engine.process(batch, trim=trim, pileup=pileup, sample=sample, first=first, n=count)
    if pileup and world > 1:
        if hasattr(engine, "ins_pack_device"):
            import torch
            ex = exchange or DeepExchange(engine, group)
            ex.run(None)
            torch.cuda.synchronize(engine.device)
            if exchange is None:
                ex.close()
        else:
            allreduce_counts(engine, group)
            allgather_insertions(engine, group)
    return t, (first, count)
