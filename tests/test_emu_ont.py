"""CPU: the warp-per-read kernel for indel-rich batches (amplipy_b200/csrc/amp_ont.cuh) run by tests/emu (every CUDA thread a
fiber) against the golden fixtures of the unmodified reference and against the oracle on seeded ONT-like, Illumina-like and
adversarial inputs: the range form of trim_read (ops[ka..kb] + clips) and the per-op pileup must be exact, everything the
closed form declines must come out of the generic path unchanged."""
import numpy as np
import pytest

import emu_driver
import golden_io
import parity
from amplipy_b200 import synth
from amplipy_b200.batch import ReadBatch

CASES = golden_io.list_cases()


def ont(**knobs):
    return lambda **kw: emu_driver.EmuEngine(kernel="ont", **knobs, **kw)


@pytest.mark.parametrize("name", CASES)
def test_aio(name):
    parity.check_case_aio(ont(), name)


@pytest.mark.parametrize("name", CASES)
def test_pileup_only(name):
    parity.check_case_pileup_only(ont(), name)


@pytest.mark.parametrize("name", ["cfg4_ont", "fuzz1", "quirks"])
def test_launch_shapes(name):
    parity.check_case_aio(ont(grid=3, warps=2, wt=64), name)
    parity.check_case_aio(ont(grid=1, warps=1), name)


def _scheme(L=29903, n_amp=98, seed=2, n_alt=0):
    g = synth.random_genome(L, 7)
    primers, amps = synth.make_scheme(L, n_amp, seed=seed, n_alt=n_alt)
    return g, [(s, e) for s, e, _ in primers], amps


@pytest.mark.parametrize("seed,mq,offset", [(61, 20, 0), (62, 10, 0), (63, 7, 2), (64, 25, 0)])
def test_ont_like_vs_oracle(oracle_lib, seed, mq, offset):
    g, prim, amps = _scheme(seed=seed % 3 + 2, n_alt=4 * (seed % 2))
    b = synth.ont_batch(g, amps, 1500, seed=seed)
    parity.check_against_oracle(ont(grid=3, warps=4), oracle_lib, b, g, prim, mq=mq, offset=offset, ins_slots=1 << 18, arena=1 << 24)


@pytest.mark.parametrize("seed", [71, 72, 73, 74, 75, 76])
def test_fuzz_vs_oracle(oracle_lib, seed):
    L = 4000
    g = synth.random_genome(L, 5)
    primers, _ = synth.make_scheme(L, 11, amp_len=350, seed=3, n_alt=3)
    prim = [(s, e) for s, e, _ in primers]
    b = ReadBatch.from_records(synth.fuzz_records(L, 2500, seed=seed, ont_like=bool(seed & 1), edges=0.1 if seed > 74 else 0.0))
    parity.check_against_oracle(ont(grid=2, warps=3), oracle_lib, b, g, prim, offset=[0, 2, 7][seed % 3], mq=[20, 0, 33][seed % 3],
                                w=[4, 4, 9][seed % 3], inc=bool(seed & 2))


def test_illumina_with_indels_vs_oracle(oracle_lib):
    g, prim, amps = _scheme(seed=3, n_alt=5)
    b = synth.illumina_batch(g, amps, 8000, seed=77, p_ins=0.3, p_del=0.3, p_clip=0.3, p_hard=0.1, p_short=0.2)
    rng = np.random.default_rng(3)
    q = b.qual
    for i in range(0, b.n, 2):
        lo, hi = int(b.qual_off[i]), int(b.qual_off[i + 1])
        if hi - lo >= 12:
            k = int(rng.integers(1, 12)); q[rng.integers(lo, hi, k)] = rng.integers(0, 20, k)
    parity.check_against_oracle(ont(grid=3, warps=4), oracle_lib, b, g, prim)
    parity.check_against_oracle(ont(grid=3, warps=4), oracle_lib, b, g, prim, mq=30, offset=3)
