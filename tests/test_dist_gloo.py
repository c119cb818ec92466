"""CPU, world_size 2 (gloo): the read-sharded multi-rank path -- contiguous read ranges per rank, one
all-reduce of the count matrix, all-gather + merge of the insertion tables -- must reproduce the
single-rank golden result exactly.  The kernels run in the emulator; the sharding / exchange code is
the product's (amplipy_b200/dist.py)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import emu_driver, golden_io, parity
    from amplipy_b200 import dist as adist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, meta, arr = golden_io.load_case(name)
    eng = emu_driver.EmuEngine(**parity.engine_kwargs(meta))
    t, (first, count) = adist.process_deep_sample(eng, b)
    # per-read outputs are valid on the owning rank
    sl = slice(first, first + count)
    assert np.array_equal(t.pos[sl], arr["t_pos"][sl]) and np.array_equal(t.flags[sl], arr["t_flags"][sl])
    assert np.array_equal(eng.counts(), arr["counts_aio"])          # every rank holds the total
    assert eng.insertions().as_dict() == golden_io.ins_from_meta(meta)
    p = meta["params"]
    ref_seq = bytes(arr["ref_seq"]).decode()
    res = eng.call(ref_seq, p["min_depth_consensus"], p["min_freq_consensus"], p["min_depth_variants"], p["min_freq_variants"])
    from amplipy_b200 import calling
    assert calling.consensus_string(res, eng.insertions(), 0, p["unknown_symbol"]) == meta["consensus"]
    dist.barrier()
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["cfg3_deep_alt", "cfg4_ont"])
def test_read_sharded_two_ranks(tmp_path, name):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, name, str(tmp_path)), nprocs=2, join=True)
    assert os.path.isfile(tmp_path / "ok0") and os.path.isfile(tmp_path / "ok1")


def test_partitions():
    from amplipy_b200 import dist as adist
    assert adist.plate_assignment(10, 1, 4) == [1, 5, 9]
    cover = []
    for r in range(3):
        f, c = adist.read_range(10, r, 3)
        cover += list(range(f, f + c))
    assert cover == list(range(10))
