"""CPU: the device-side BGZF / BAM decoder (amplipy_b200/csrc/amp_bgzf.cuh), run by tests/emu with one warp of fibers per
block, against zlib and the host decoder: byte-equal output on dynamic, fixed and stored deflate blocks, every compression
level, long and overlapping matches; malformed streams end in an error code, never in an out-of-bounds write."""
import ctypes
import zlib

import numpy as np
import pytest

import emu_driver
from amplipy_b200 import alnio, synth


def _inflate(raw, out_len, slack=64):
    lib = emu_driver.lib()
    src = np.zeros(len(raw) + 8, np.uint8)
    src[:len(raw)] = np.frombuffer(raw, np.uint8)
    out = np.full(out_len + slack, 0xAB, np.uint8)
    err = lib.emu_inflate(ctypes.c_void_p(src.ctypes.data), ctypes.c_longlong(len(raw)), ctypes.c_void_p(out.ctypes.data),
                          ctypes.c_longlong(out_len))
    assert (out[out_len:] == 0xAB).all(), "wrote past the end of the output"
    return err, out[:out_len].tobytes()


def _deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    return c.compress(data) + c.flush()


def _payloads():
    rng = np.random.default_rng(5)
    L = 3000
    g = synth.random_genome(L, 3)
    _, amps = synth.make_scheme(L, 8, seed=2)
    b = synth.illumina_batch(g, amps, 150, seed=6)
    bam_like = b.qual.tobytes()[:20000] + b.seq.tobytes()[:8000] + b.cigar.tobytes() + b.pos.tobytes()
    return {
        "bam_like": bam_like,
        "text": (b"@SQ\tSN:ref\tLN:29903\n" * 500)[:9000],
        "random": rng.integers(0, 256, 5000, dtype=np.uint8).tobytes(),
        "zeros": bytes(40000),
        "rle": b"ab" * 7000 + b"x" * 300 + b"abc" * 999,
        "tiny": b"A",
        "empty": b"",
        "full_block": rng.integers(0, 4, 65280, dtype=np.uint8).tobytes(),
    }


@pytest.mark.parametrize("name", sorted(_payloads()))
@pytest.mark.parametrize("level", [1, 6, 9])
def test_inflate_matches_zlib(name, level):
    data = _payloads()[name]
    err, out = _inflate(_deflate(data, level), len(data))
    assert err == 0 and out == data


def test_fixed_and_stored_blocks():
    data = _payloads()["bam_like"][:3000]
    for raw in (_deflate(data, 6, zlib.Z_FIXED), _deflate(data, 0), _deflate(b"hello hello hello", 9, zlib.Z_FIXED)):
        want = zlib.decompress(raw, -15)
        err, out = _inflate(raw, len(want))
        assert err == 0 and out == want


def test_several_deflate_blocks_in_one_stream():
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    parts = [_payloads()["text"], _payloads()["random"][:2000], _payloads()["rle"][:5000]]
    raw = b"".join(c.compress(p) + c.flush(zlib.Z_FULL_FLUSH) for p in parts) + c.flush()
    want = b"".join(parts)
    err, out = _inflate(raw, len(want))
    assert err == 0 and out == want


def test_malformed_streams_raise_error_codes():
    data = _payloads()["bam_like"][:6000]
    raw = _deflate(data)
    err, _ = _inflate(raw, len(data) - 10)            # ISIZE too small: must stop at the end of the output
    assert err != 0
    err, _ = _inflate(raw, len(data) + 10)            # ISIZE too large
    assert err != 0
    err, _ = _inflate(raw[:len(raw) // 2], len(data))  # truncated
    assert err != 0
    rng = np.random.default_rng(1)
    for k in range(40):                                # bit flips: any outcome but a crash / overrun; correct output only if err == 0
        bad = bytearray(raw)
        i = int(rng.integers(0, len(bad)))
        bad[i] ^= 1 << int(rng.integers(0, 8))
        err, out = _inflate(bytes(bad), len(data))
        if err == 0:
            try:
                assert zlib.decompress(bytes(bad), -15) == out
            except zlib.error:
                pass                                    # zlib is stricter about some malformed codes


def test_bam_chain_and_scatter_match_the_host_decoder(tmp_path):
    """Block-aligned BAM (records never straddle BGZF blocks, as htslib writes them): per-block totals + scatter == amp_bam_fill."""
    L = 3000
    g = synth.random_genome(L, 3)
    _, amps = synth.make_scheme(L, 8, seed=2)
    b = synth.illumina_batch(g, amps, 700, seed=8, p_ins=0.2, p_del=0.2)
    path = str(tmp_path / "a.bam")
    alnio.write_bam(path, "@HD\tVN:1.6\n@SQ\tSN:ref\tLN:%d\n@PG\tID:x\tPN:x\n" % L, [("ref", L)], b)
    raw_file = open(path, "rb").read()
    a = alnio._read_bam(raw_file)
    info = alnio.bgzf_blocks(raw_file)
    payload = a.bam_buf
    lib = emu_driver.lib()
    body = int(a.bam_rec_off[0])
    starts = np.concatenate([[0], np.cumsum(info["out_len"])]).astype(np.int64)
    n = b.n
    got = dict(pos=np.zeros(n, np.int32), flag=np.zeros(n, np.uint16), tlen=np.zeros(n, np.int32), cig_off=np.zeros(n + 1, np.uint32),
               cigar=np.zeros(b.cigar.size, np.uint32), seq_off=np.zeros(n + 1, np.uint32), seq=np.zeros(b.seq.size, np.uint8),
               qual_off=np.zeros(n + 1, np.uint32), qual=np.zeros(b.qual.size, np.uint8), rec_off=np.zeros(n, np.uint64))
    r0 = c0 = s0 = q0 = 0
    p = lambda x: ctypes.c_void_p(x.ctypes.data)
    for k in range(len(starts) - 1):
        lo, hi = max(int(starts[k]), body), int(starts[k + 1])
        if lo >= hi:
            continue
        tot = (ctypes.c_ulonglong * 4)()
        assert lib.emu_bam_totals(p(payload), ctypes.c_longlong(lo), ctypes.c_longlong(hi), tot) == 1, "record straddles block %d" % k
        nr, nc, ns, nq = (int(x) for x in tot)
        lib.emu_bam_scatter(p(payload), ctypes.c_longlong(lo), ctypes.c_longlong(hi), p(got["pos"][r0:]), p(got["flag"][r0:]),
                            p(got["tlen"][r0:]), p(got["cig_off"][r0:]), p(got["cigar"][c0:]), p(got["seq_off"][r0:]), p(got["seq"][s0:]),
                            p(got["qual_off"][r0:]), p(got["qual"][q0:]), p(got["rec_off"][r0:]))
        got["cig_off"][r0:r0 + nr] += c0; got["seq_off"][r0:r0 + nr] += s0; got["qual_off"][r0:r0 + nr] += q0
        r0 += nr; c0 += nc; s0 += ns; q0 += nq
    assert r0 == n
    got["cig_off"][n] = c0; got["seq_off"][n] = s0; got["qual_off"][n] = q0
    for f in ("pos", "flag", "tlen", "cig_off", "cigar", "seq_off", "seq", "qual_off", "qual"):
        assert np.array_equal(got[f], getattr(a.batch, f)), f
    assert np.array_equal(got["rec_off"].astype(np.int64), a.bam_rec_off)


def test_record_rewrite_matches_the_host_writer(oracle_lib, tmp_path):
    """bam_rewrite_record (what amp_decoded_write_bam runs a warp per kept read) against amp_bam_rewrite of the host codec: trimmed
    positions / CIGARs from the oracle patched into the records of a BAM stream, byte for byte."""
    import os
    from oracle import oracle
    from amplipy_b200.primers import find_overlapping_primers, max_primer_len
    L = 3000
    g = synth.random_genome(L, 3)
    primers, amps = synth.make_scheme(L, 8, seed=2)
    prim = [(s, e) for s, e, _ in primers]
    lib = emu_driver.lib()
    lib.emu_bam_rewrite.restype = ctypes.c_longlong
    host = alnio.hostio()
    for b in (synth.illumina_batch(g, amps, 600, seed=8, p_ins=0.2, p_del=0.2, p_clip=0.3, p_hard=0.2), synth.ont_batch(g, amps, 60, seed=9)):
        path = os.path.join(str(tmp_path), "x%d.bam" % b.n)
        alnio.write_bam(path, "@HD\tVN:1.6\n@SQ\tSN:ref\tLN:%d\n@PG\tID:s\tPN:s\n" % L, [("ref", L)], b)
        a = alnio._read_bam(open(path, "rb").read())
        mn, mx = oracle.find_overlapping_primers(L, prim, 0)
        t = oracle.trim_batch(b, L, mn, mx, max_primer_len(prim))
        keep = (t["flags"] & 8) != 0
        sel = np.flatnonzero(keep).astype(np.int64)
        assert 0 < sel.size < b.n
        p = lambda x: ctypes.c_void_p(x.ctypes.data)
        pos, ncig, cig = np.ascontiguousarray(t["pos"], np.int32), np.ascontiguousarray(t["ncig"], np.uint16), np.ascontiguousarray(t["cigar"], np.uint32)
        size = host.amp_bam_rewrite(p(a.bam_buf), p(a.bam_rec_off), p(sel), ctypes.c_longlong(sel.size), p(pos), p(ncig), p(b.cig_off), p(cig), None)
        want = np.zeros(size + 8, np.uint8)
        host.amp_bam_rewrite(p(a.bam_buf), p(a.bam_rec_off), p(sel), ctypes.c_longlong(sel.size), p(pos), p(ncig), p(b.cig_off), p(cig), p(want))
        rec_off = a.bam_rec_off.astype(np.uint64)
        got = np.full(size + 8, 0xEE, np.uint8)
        n = lib.emu_bam_rewrite(p(a.bam_buf), p(rec_off), p(sel), ctypes.c_longlong(sel.size), p(pos), p(ncig), p(b.cig_off), p(cig), p(got))
        assert n == size and np.array_equal(got[:size], want[:size]) and (got[size:] == 0xEE).all()
