"""CPU: the CUDA kernels' own source (amplipy_b200/csrc/*.cuh), compiled for the host by tests/emu,
must reproduce the golden fixtures produced by the unmodified reference."""
import pytest

import emu_driver
import golden_io
import parity

CASES = golden_io.list_cases()


@pytest.mark.parametrize("name", CASES)
def test_aio(name):
    parity.check_case_aio(emu_driver.EmuEngine, name)


@pytest.mark.parametrize("name", CASES)
def test_pileup_only(name):
    parity.check_case_pileup_only(emu_driver.EmuEngine, name)


@pytest.mark.parametrize("name", ["cfg1_example", "cfg2_illumina", "fuzz1"])
def test_trim_then_variants_pipeline(name):
    parity.check_case_pipeline(emu_driver.EmuEngine, name)


@pytest.mark.parametrize("knobs", [dict(grid=1, reads_per_tile=7), dict(grid=3, reads_per_tile=64, maxseg=16),
                                   dict(grid=5, reads_per_tile=33, wt=32), dict(grid=2, reads_per_tile=100, qbytes=256),
                                   dict(threads=64, reads_per_tile=256)])
@pytest.mark.parametrize("name", ["cfg2_illumina", "cfg4_ont", "fuzz3"])
def test_tile_shapes(name, knobs):
    """Overflowing run lists, windows narrower than a read, staging buffers smaller than a tile, and
    tiles larger than the CTA must all take their exact slow paths."""
    parity.check_case_aio(lambda **kw: emu_driver.EmuEngine(**kw, **knobs), name)


def test_long_reads_emulated(oracle_lib):
    """Rows larger than the staging buffers and CIGARs longer than the per-thread arrays (global scratch rows)."""
    import numpy as np
    from test_gpu_parity import _long_read_batch
    from amplipy_b200 import synth
    from amplipy_b200.primers import find_overlapping_primers, max_primer_len
    L = 12000
    g, b = _long_read_batch(L, 40, seed=5, read_len=(1500, 4000))
    primers, _ = synth.make_scheme(L, 30, amp_len=400, seed=9)
    prim = [(s, e) for s, e, _ in primers]
    mn, mx = oracle_lib.find_overlapping_primers(L, prim, 0)
    want = oracle_lib.trim_batch(b, L, mn, mx, max_primer_len(prim), 15)
    wc, wins, nerr = oracle_lib.pileup_batch(b, L, 15, trimmed=want)
    eng = emu_driver.EmuEngine(ref_len=L, primer_tables=find_overlapping_primers(L, prim, 0), max_primer_len=max_primer_len(prim),
                               min_quality=15, ins_slots=1 << 18, ins_arena_bytes=1 << 24)
    t = eng.process(b)
    assert eng.error_flags() == 0 and nerr == 0
    assert np.array_equal(t.pos, want["pos"]) and np.array_equal(t.flags, want["flags"])
    assert np.array_equal(t.ncig.astype(np.int32), want["ncig"])
    assert np.array_equal(eng.counts().astype(np.int64), wc)
    assert eng.insertions().as_dict() == wins
