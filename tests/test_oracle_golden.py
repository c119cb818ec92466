"""CPU: the C oracle (oracle/amplipy_oracle.c) must reproduce every golden fixture, i.e. the outputs
of the unmodified reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

import golden_io

CASES = golden_io.list_cases()


def _tables(oracle, meta):
    p = meta["params"]
    return oracle.find_overlapping_primers(meta["L"], [tuple(x) for x in meta["primers"]], p["offset"])


@pytest.mark.parametrize("name", CASES)
def test_primer_tables(oracle_lib, name):
    _, meta, arr = golden_io.load_case(name)
    mn, mx = _tables(oracle_lib, meta)
    assert np.array_equal(mn, arr["min_primer_start"])
    assert np.array_equal(mx, arr["max_primer_end"])


@pytest.mark.parametrize("name", CASES)
def test_trim(oracle_lib, name):
    b, meta, arr = golden_io.load_case(name)
    p = meta["params"]
    mn, mx = _tables(oracle_lib, meta)
    mpl = max(e - s for s, e in meta["primers"])
    t = oracle_lib.trim_batch(b, meta["L"], mn, mx, mpl, p["min_quality"], p["window"], p["min_length"],
                              p["include_no_primer"])
    assert np.array_equal(t["flags"], arr["t_flags"])
    assert np.array_equal(t["pos"], arr["t_pos"])
    assert np.array_equal(t["ncig"], arr["t_ncig"])
    assert np.array_equal(t["cigar"], arr["t_cigar"])


@pytest.mark.parametrize("name", CASES)
def test_pileup_and_call(oracle_lib, name):
    b, meta, arr = golden_io.load_case(name)
    p = meta["params"]
    L = meta["L"]
    t = {"pos": arr["t_pos"], "ncig": arr["t_ncig"], "cigar": arr["t_cigar"], "flags": arr["t_flags"]}
    counts, ins, nerr = oracle_lib.pileup_batch(b, L, p["min_quality"], trimmed=t)
    assert nerr == 0
    assert np.array_equal(counts, arr["counts_aio"])
    assert ins == golden_io.ins_from_meta(meta)
    # untrimmed pileup (variants/consensus subcommands on arbitrary input)
    craw, iraw, _ = oracle_lib.pileup_batch(b, L, p["min_quality"])
    assert np.array_equal(craw, arr["counts_raw"])
    assert iraw == {(int(a), s): int(c) for a, s, c in meta["insertions_raw"]}
    ref_seq = bytes(arr["ref_seq"]).decode()
    res = oracle_lib.call(counts, ins, ref_seq, True, p["min_depth_consensus"], p["min_freq_consensus"], True,
                          p["min_depth_variants"], p["min_freq_variants"])
    assert np.array_equal(res["depth"], arr["depth_aio"])
    assert np.array_equal(res["al_off"], arr["al_off"])
    n = int(res["al_off"][-1])
    assert np.array_equal(res["al_count"][:n], arr["al_count"])
    assert np.array_equal(res["al_freq"][:n], arr["al_freq"])       # float64, bit-exact
    assert [res["sym"](int(s)) for s in res["al_sym"][:n]] == meta["al_sym"]
    assert oracle_lib.consensus_string(res, p["unknown_symbol"]) == meta["consensus"]
    got = [[v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], list(v[8])] for v in oracle_lib.variant_records(res, ref_seq)]
    assert got == meta["variants"]


def test_kept_only_pileup(oracle_lib):
    """trim -> variants pipeline: reads dropped by the write gate are not piled up (AmpliPy.py:910-911)."""
    b, meta, arr = golden_io.load_case("cfg2_illumina")
    flags = arr["t_flags"].copy()
    flags[(flags & 8) == 0] |= 16           # mark dropped reads as skipped for the second step
    t = {"pos": arr["t_pos"], "ncig": arr["t_ncig"], "cigar": arr["t_cigar"], "flags": flags}
    counts, ins, _ = oracle_lib.pileup_batch(b, meta["L"], meta["params"]["min_quality"], trimmed=t)
    assert np.array_equal(counts, arr["counts_kept"])
    assert ins == {(int(a), s): int(c) for a, s, c in meta["insertions_kept"]}
