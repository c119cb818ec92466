"""CPU: the warp-autonomous kernel (amplipy_b200/csrc/amp_warp.cuh) run by tests/emu with every CUDA thread as a
fiber (shuffles / ballots / reductions are rendezvous points) must reproduce the golden fixtures produced by the
unmodified reference, and the oracle on larger seeded inputs."""
import numpy as np
import pytest

import emu_driver
import golden_io
import parity
from amplipy_b200 import synth

CASES = golden_io.list_cases()


def v7(**knobs):
    return lambda **kw: emu_driver.EmuEngine(kernel="v7", **knobs, **kw)


@pytest.mark.parametrize("name", CASES)
def test_aio(name):
    parity.check_case_aio(v7(), name)


@pytest.mark.parametrize("name", CASES)
def test_pileup_only(name):
    parity.check_case_pileup_only(v7(), name)


@pytest.mark.parametrize("name", ["cfg1_example", "cfg2_illumina", "fuzz1"])
def test_trim_then_variants_pipeline(name):
    parity.check_case_pipeline(v7(), name)


@pytest.mark.parametrize("knobs", [dict(warps=2, grid=3, batch_reads=5, wt=32), dict(warps=1, grid=1, batch_reads=32, wt=64),
                                   dict(warps=3, grid=7, batch_reads=17), dict(warps=4, grid=2, batch_reads=1)])
@pytest.mark.parametrize("name", ["cfg2_illumina", "cfg4_ont", "fuzz3", "quirks"])
def test_launch_shapes(name, knobs):
    """Windows narrower than a read (global-atomic path), partial batches, few warps draining long queues."""
    parity.check_case_aio(v7(**knobs), name)


def _scheme(L=29903, n_amp=98, seed=2, n_alt=0):
    g = synth.random_genome(L, 7)
    primers, amps = synth.make_scheme(L, n_amp, seed=seed, n_alt=n_alt)
    return g, [(s, e) for s, e, _ in primers], amps


def test_illumina_vs_oracle(oracle_lib):
    g, prim, amps = _scheme()
    b = synth.illumina_batch(g, amps, 40_000, seed=31, snvs=[(1000, "T", 0.5), (20000, "A", 0.03)])
    emu_driver.v7_stats()
    parity.check_against_oracle(v7(), oracle_lib, b, g, prim)
    fast, generic = emu_driver.v7_stats()
    assert fast + generic == b.n and fast > 0.85 * b.n, (fast, generic)   # the cooperative path is the one being tested


@pytest.mark.parametrize("mq,w,offset,inc", [(30, 4, 0, False), (11, 4, 3, True), (0, 4, 0, False), (20, 6, 0, False),
                                             (127, 4, 0, True), (200, 4, 0, True)])
def test_illumina_parameter_grid_vs_oracle(oracle_lib, mq, w, offset, inc):
    """Other thresholds; window widths != 4 and thresholds outside the SIMD compare send every read down the generic path."""
    g, prim, amps = _scheme(seed=3, n_alt=10)
    b = synth.illumina_batch(g, amps, 6_000, seed=32 + mq)
    parity.check_against_oracle(v7(grid=4, warps=4), oracle_lib, b, g, prim, offset=offset, mq=mq, w=w, inc=inc)


def test_unsorted_vs_oracle(oracle_lib):
    g, prim, amps = _scheme()
    b = synth.illumina_batch(g, amps, 12_000, seed=33, sort=False)
    parity.check_against_oracle(v7(grid=5, warps=4), oracle_lib, b, g, prim)


def test_longer_reads_vs_oracle(oracle_lib):
    """250-bp reads: fewer reads per batch, aligned runs that use all eight words per lane, some beyond them."""
    g, prim, amps = _scheme(seed=5)
    b = synth.illumina_batch(g, amps, 8_000, seed=34, read_len=250)
    parity.check_against_oracle(v7(grid=3, warps=4), oracle_lib, b, g, prim)
    b = synth.illumina_batch(g, amps, 4_000, seed=35, read_len=301)
    parity.check_against_oracle(v7(grid=3, warps=4), oracle_lib, b, g, prim)


def test_window_edges_vs_oracle(oracle_lib):
    """Dense low-quality stretches so that failing windows land on every alignment of the run and in the shrinking
    windows at its open end, on both strands."""
    g, prim, amps = _scheme(seed=6)
    rng = np.random.default_rng(7)
    b = synth.illumina_batch(g, amps, 10_000, seed=36)
    q = b.qual.copy()
    for i in range(b.n):
        lo, hi = int(b.qual_off[i]), int(b.qual_off[i + 1])
        kind = rng.integers(0, 6)
        if kind == 0:      # one weak window at a random place
            p = rng.integers(lo, hi - 4); q[p:p + 4] = rng.integers(0, 25, 4)
        elif kind == 1:    # weak last / first bases only
            k = rng.integers(1, 4); q[hi - k:hi] = rng.integers(0, 20, k)
        elif kind == 2:
            k = rng.integers(1, 4); q[lo:lo + k] = rng.integers(0, 20, k)
        elif kind == 3:    # sums right at the threshold
            p = rng.integers(lo, hi - 4); q[p:p + 4] = [20, 20, 20, rng.integers(19, 22)]
        elif kind == 4:
            q[lo:hi] = rng.integers(15, 26, hi - lo)
    b.qual[:] = q
    emu_driver.v7_stats()
    parity.check_against_oracle(v7(grid=3, warps=4), oracle_lib, b, g, prim)
    fast, generic = emu_driver.v7_stats()
    assert fast > 0.85 * b.n, (fast, generic)


def test_register_classification_matches_array_form():
    """classify_simple3 (CIGAR words in registers, used by the pipelined phase A) == classify_simple on random CIGARs."""
    import ctypes
    lib = emu_driver.lib()
    rng = np.random.default_rng(11)
    n_simple = 0
    for _ in range(20000):
        nc = int(rng.integers(0, 6))
        ops = rng.choice([0, 0, 0, 4, 4, 1, 2, 5, 7, 8], size=nc)
        lens = rng.integers(0, 60, size=nc)
        cig = (lens.astype(np.uint32) << 4 | ops.astype(np.uint32)).astype(np.uint32)
        buf = np.zeros(8, np.uint32); buf[:nc] = cig
        q_len = int(sum(l for l, o in zip(lens, ops) if o in (0, 1, 4, 7, 8)))
        for l_seq in (q_len, q_len + 1):
            assert lib.emu_classify_agree(ctypes.c_void_p(buf.ctypes.data), nc, l_seq) == 1, (ops, lens, l_seq)
        n_simple += nc >= 1 and all(o in (0, 4, 7, 8) for o in ops)
    assert n_simple > 1000


def _weaken_qualities(b, seed):
    """Weak windows / ends / whole rows at random places: quality clips that stop in front of, inside and beyond an indel."""
    rng = np.random.default_rng(seed)
    q = b.qual.copy()
    for i in range(b.n):
        lo, hi = int(b.qual_off[i]), int(b.qual_off[i + 1])
        if hi - lo < 12:
            continue
        kind = rng.integers(0, 8)
        if kind == 0:
            p = rng.integers(lo, hi - 4); q[p:p + 4] = rng.integers(0, 25, 4)
        elif kind == 1:
            k = rng.integers(1, 4); q[hi - k:hi] = rng.integers(0, 20, k)
        elif kind == 2:
            k = rng.integers(1, 4); q[lo:lo + k] = rng.integers(0, 20, k)
        elif kind == 3:
            p = rng.integers(lo, hi - 4); q[p:p + 4] = [20, 20, 20, rng.integers(19, 22)]
        elif kind == 4:
            q[lo:hi] = rng.integers(15, 26, hi - lo)
        elif kind == 5:    # single weak bases sprinkled over the row (inserted bases below the threshold, split insertions)
            k = rng.integers(1, 12); q[rng.integers(lo, hi, k)] = rng.integers(0, 20, k)
        elif kind == 6:    # everything weak
            q[lo:hi] = rng.integers(0, 19, hi - lo)
    b.qual[:] = q
    return b


@pytest.mark.parametrize("seed,mq", [(41, 20), (42, 20), (43, 30), (44, 5)])
def test_single_indel_shapes_vs_oracle(oracle_lib, seed, mq):
    """[H][S] M (I|D) M [S][H] reads are finished in registers like [S]M[S]: heavy indel / clip rates and quality patterns
    that move the quality clip across the indel on both strands; they must stay on the cooperative path."""
    g, prim, amps = _scheme(seed=seed % 3 + 2, n_alt=5 * (seed % 2))
    b = synth.illumina_batch(g, amps, 12_000, seed=seed, p_ins=0.3, p_del=0.3, p_clip=0.3, p_hard=0.1, p_short=0.2)
    b = _weaken_qualities(b, seed + 100)
    emu_driver.v7_stats()
    parity.check_against_oracle(v7(grid=3, warps=4), oracle_lib, b, g, prim, mq=mq)
    fast, generic = emu_driver.v7_stats()
    assert fast > 0.8 * b.n, (fast, generic)


def test_many_generic_reads_overflow_the_shared_list(oracle_lib):
    """Window width 6 sends every read down the generic path: the shared-memory list overflows into P.glist."""
    g, prim, amps = _scheme(seed=3)
    b = synth.illumina_batch(g, amps, 5_000, seed=45, p_ins=0.2, p_del=0.2)
    parity.check_against_oracle(v7(grid=2, warps=3), oracle_lib, b, g, prim, w=6)


@pytest.mark.parametrize("kernel", ["v7", "tile"])
def test_reads_at_the_ends_of_the_reference_vs_oracle(oracle_lib, kernel):
    """Reads at position 0 / ending on the last base, a primer that reaches the last base: a read swallowed by the start clip
    ends up at pos == L with an all-S CIGAR; the reference piles it up without touching anything and so must the kernels."""
    from amplipy_b200.batch import ReadBatch
    L = 240
    g = synth.random_genome(L, 3)
    prim = [(0, 22), (60, 84), (150, 171), (L - 24, L)]
    recs = synth.fuzz_records(L, 1500, seed=111, max_len=40, edges=0.5)
    recs += [(L - 2, 0, 0, [(0, 2)], "AC", [30, 30]), (L - 6, 16, 0, [(0, 6)], "ACGTAC", [30] * 6)]
    b = ReadBatch.from_records(sorted(recs, key=lambda r: r[0]))
    mk = (lambda **kw: emu_driver.EmuEngine(kernel=kernel, grid=2, **({"warps": 3} if kernel == "v7" else {}), **kw))
    parity.check_against_oracle(mk, oracle_lib, b, g, prim, offset=1, ml=5)
