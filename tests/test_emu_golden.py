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
