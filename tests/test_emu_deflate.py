"""CPU: the device-side BGZF compressor (amplipy_b200/csrc/amp_deflate.cuh), run by tests/emu with one warp of fibers per block:
whatever it writes must inflate (zlib, and the repo's own device-side inflater) to the input, and its CRC-32 must be zlib's."""
import ctypes
import zlib

import numpy as np
import pytest

import emu_driver
from amplipy_b200 import synth


def _deflate(data, cap_words=None):
    lib = emu_driver.lib()
    n = len(data)
    src = np.zeros(n + 16, np.uint8)
    src[:n] = np.frombuffer(data, np.uint8)
    cap = (n + 3) // 4 if cap_words is None else cap_words
    out = np.full(cap + 64, 0xABABABAB, np.uint32)
    crc = ctypes.c_uint32(0)
    lib.emu_deflate.restype = ctypes.c_int
    r = lib.emu_deflate(ctypes.c_void_p(src.ctypes.data), ctypes.c_int(n), ctypes.c_void_p(out.ctypes.data), ctypes.c_int(cap), ctypes.byref(crc))
    return r, out.view(np.uint8)[:max(r, 0)].tobytes(), crc.value


def _payloads():
    rng = np.random.default_rng(11)
    L = 3000
    g = synth.random_genome(L, 3)
    _, amps = synth.make_scheme(L, 8, seed=2)
    b = synth.illumina_batch(g, amps, 200, seed=6)
    o = synth.ont_batch(g, amps, 30, seed=7)
    return {
        "bam_like": b.qual.tobytes()[:30000] + b.seq.tobytes()[:12000] + b.cigar.tobytes() + b.pos.tobytes(),
        "ont_like": (o.qual.tobytes() + o.seq.tobytes() + o.cigar.tobytes())[:65280],
        "text": (b"@SQ\tSN:ref\tLN:29903\n" * 500)[:9000],
        "zeros": bytes(40000),
        "rle": b"ab" * 7000 + b"x" * 300 + b"abc" * 999 + b"q" * 5,
        "low_entropy": rng.integers(0, 4, 65280, dtype=np.uint8).tobytes(),
        "far_matches": (lambda a: a + bytes(32100) + a + bytes(300) + a[:300] + bytes(31000) + a)(rng.integers(0, 256, 600, dtype=np.uint8).tobytes()),
        "beyond_window": rng.integers(0, 256, 33000, dtype=np.uint8).tobytes() + rng.integers(0, 256, 200, dtype=np.uint8).tobytes() * 100,
        "short": b"ACGTACGTACGTACGTAC",
        "odd_tail": b"x" * 4097 + b"yz",
    }


@pytest.mark.parametrize("name", sorted(_payloads()))
def test_round_trip_through_zlib(name):
    data = _payloads()[name]
    r, comp, crc = _deflate(data, cap_words=None if len(data) > 1000 else len(data) // 4 + 64)   # (the kernel stores blocks that small)
    assert crc == zlib.crc32(data)
    assert r > 0, "compressible data must come back compressed"
    assert r < len(data) or len(data) < 1000
    d = zlib.decompressobj(-15)
    got = d.decompress(comp) + d.flush()
    assert got == data and d.eof and d.unused_data == b""


@pytest.mark.parametrize("offset", [0, 1, 2, 3])
def test_any_alignment_of_the_input(offset):
    data = _payloads()["bam_like"][offset:offset + 20011]
    lib = emu_driver.lib()
    n = len(data)
    src = np.zeros(n + 32, np.uint8)
    src[offset:offset + n] = np.frombuffer(data, np.uint8)
    out = np.zeros((n + 3) // 4 + 64, np.uint32)
    crc = ctypes.c_uint32(0)
    lib.emu_deflate.restype = ctypes.c_int
    r = lib.emu_deflate(ctypes.c_void_p(src.ctypes.data + offset), ctypes.c_int(n), ctypes.c_void_p(out.ctypes.data), ctypes.c_int((n + 3) // 4),
                        ctypes.byref(crc))
    assert r > 0 and crc.value == zlib.crc32(data)
    assert zlib.decompress(out.view(np.uint8)[:r].tobytes(), -15) == data


def test_incompressible_data_is_declined_without_overrunning_the_slot():
    rng = np.random.default_rng(3)
    data = rng.integers(0, 256, 65280, dtype=np.uint8).tobytes()
    lib = emu_driver.lib()
    src = np.zeros(len(data) + 16, np.uint8); src[:len(data)] = np.frombuffer(data, np.uint8)
    cap = (len(data) + 3) // 4
    out = np.full(cap + 64, 0xABABABAB, np.uint32)
    crc = ctypes.c_uint32(0)
    lib.emu_deflate.restype = ctypes.c_int
    r = lib.emu_deflate(ctypes.c_void_p(src.ctypes.data), ctypes.c_int(len(data)), ctypes.c_void_p(out.ctypes.data), ctypes.c_int(cap), ctypes.byref(crc))
    assert r == -1 and crc.value == zlib.crc32(data)
    assert (out[cap + 41:] == 0xABABABAB).all()


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 31, 32, 33, 2047, 2048, 2049, 4096, 65279, 65280])
def test_crc_and_sizes(n):
    rng = np.random.default_rng(n)
    data = rng.integers(65, 69, n, dtype=np.uint8).tobytes()
    r, comp, crc = _deflate(data, cap_words=(n + 3) // 4 + 64)
    assert crc == zlib.crc32(data)
    if n >= 16:
        assert r > 0 and zlib.decompress(comp, -15) == data
    else:
        assert r == -1


def test_output_feeds_the_device_inflater():
    data = _payloads()["bam_like"]
    r, comp, _ = _deflate(data)
    lib = emu_driver.lib()
    src = np.zeros(len(comp) + 8, np.uint8); src[:len(comp)] = np.frombuffer(comp, np.uint8)
    out = np.zeros(len(data) + 64, np.uint8)
    err = lib.emu_inflate(ctypes.c_void_p(src.ctypes.data), ctypes.c_longlong(len(comp)), ctypes.c_void_p(out.ctypes.data), ctypes.c_longlong(len(data)))
    assert err == 0 and out[:len(data)].tobytes() == data


def test_skewed_and_degenerate_alphabets():
    """Code construction corner cases: Fibonacci-like literal frequencies (an unlimited Huffman code would be 17 bits deep: the
    frequencies are halved until 15 suffice), literals only (no distance code in use: two are forced), one literal repeated (runs),
    every byte value once, and a stream long enough for several deflate blocks with different statistics."""
    rng = np.random.default_rng(23)
    fib = [1, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987, 1597, 2584]
    sym = np.repeat(np.arange(len(fib), dtype=np.uint8) * 7 + 3, fib)
    for _ in range(50):                                    # break up the runs of the frequent symbols as well as chance allows
        rng.shuffle(sym)
    cases = {
        "fibonacci": sym.tobytes(),
        "literals_only": bytes(rng.permutation(256).astype(np.uint8)) * 1 + bytes(rng.permutation(256).astype(np.uint8))[::-1][:200],
        "one_symbol": b"Q" * 5000,
        "two_symbols": (b"A" + b"B" * 3) * 700,
        "phases": bytes(rng.integers(0, 4, 30000, dtype=np.uint8)) + bytes(rng.integers(100, 228, 20000, dtype=np.uint8)) + b"xyz" * 5000,
    }
    for name, data in cases.items():
        r, comp, crc = _deflate(data, cap_words=len(data) // 4 + 200)
        assert crc == zlib.crc32(data), name
        assert r > 0, name
        d = zlib.decompressobj(-15)
        assert d.decompress(comp) + d.flush() == data and d.eof, name


def _huffman_cost(freq):
    """(total bits, depth) of an optimal unlimited prefix code (the deepest among the optimal ones: ties merge the shallower trees last)"""
    import heapq
    h = [(int(f), 0) for f in freq if f]
    if len(h) < 2:
        return sum(f for f, _ in h), 1
    heapq.heapify(h)
    cost = 0
    while len(h) > 1:
        (a, da), (b, db) = heapq.heappop(h), heapq.heappop(h)
        cost += a + b
        heapq.heappush(h, (a + b, max(da, db) + 1))
    return cost, h[0][1]


@pytest.mark.parametrize("seed", range(6))
def test_code_lengths_are_optimal_complete_and_limited(seed):
    """huff_lengths: a complete prefix code (Kraft sum exactly 1), optimal whenever the optimum fits the limit, never longer than
    the limit otherwise (Fibonacci frequencies: depth 17 unlimited), at least two symbols coded."""
    lib = emu_driver.lib()
    rng = np.random.default_rng(seed)
    fib = [1, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987, 1597, 2584]
    cases = [(rng.integers(0, 50, 286), 15), (rng.integers(0, 3, 286) * rng.integers(0, 3000, 286), 15), (np.array(fib + [0] * 268), 15),
             (np.array(fib[:12] + [0] * 7), 7), (rng.integers(0, 40, 19), 7), (np.array([0] * 30), 15), (np.array([0] * 7 + [9] + [0] * 22), 15),
             (rng.integers(0, 2000, 30), 15), (np.array([1] * 286), 15)]
    for freq, limit in cases:
        f = np.ascontiguousarray(freq, np.uint16)
        n = f.size
        out = np.full(n + 8, 0xAB, np.uint8)
        lib.emu_huff_lengths(ctypes.c_void_p(f.ctypes.data), n, limit, ctypes.c_void_p(out.ctypes.data))
        ln = out[:n].astype(int)
        assert (out[n:] == 0xAB).all()
        used = ln > 0
        assert used.sum() >= 2 and ((f > 0) <= used).all() and ln.max() <= limit
        assert sum(2.0 ** -l for l in ln[used]) == 1.0
        cost = int((f.astype(int) * ln).sum())
        best, depth = _huffman_cost(f)                      # (f: with the forced entries, if any)
        assert cost >= best
        if depth <= limit:                                  # where an unlimited optimum fits, the result must be optimal too
            assert cost == best, (cost, best, limit, depth)
        else:
            assert cost <= 1.05 * best + 8                  # the halving heuristic stays close
