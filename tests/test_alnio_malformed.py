"""CPU: malformed BAM / BGZF input must end in InputError (the reference's error convention: message + exit 1), never in a
crash: record headers whose lengths do not fit their record, BGZF extra fields that run past the block, truncated files."""
import os
import struct

import numpy as np
import pytest

from amplipy_b200 import alnio, synth
from amplipy_b200.primers import InputError


def _bam_bytes(tmp_path, n=50):
    L = 2000
    g = synth.random_genome(L, 3)
    _, amps = synth.make_scheme(L, 6, seed=2)
    b = synth.illumina_batch(g, amps, n, seed=5)
    path = os.path.join(tmp_path, "in.bam")
    alnio.write_bam(path, "@HD\tVN:1.6\n@SQ\tSN:ref\tLN:%d\n@PG\tID:x\tPN:x\n" % L, [("ref", L)], b)
    return open(path, "rb").read(), b


def _recompress(payload):
    return alnio.bgzf_compress(payload)


def _first_record_offset(payload):
    l_text = struct.unpack_from("<i", payload, 4)[0]
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", payload, p)[0]; p += 4
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", payload, p)[0]; p += 4 + l_name + 4
    return p


def test_well_formed_roundtrip(tmp_path):
    raw, b = _bam_bytes(str(tmp_path))
    a = alnio._read_bam(raw)
    assert a.batch.n == b.n and np.array_equal(a.batch.qual, b.qual) and np.array_equal(a.batch.cigar, b.cigar)


@pytest.mark.parametrize("field,value", [("l_seq", 200_000_000), ("l_seq", -5), ("n_cigar", 60000), ("l_read_name", 255)])
def test_record_lengths_beyond_the_record(tmp_path, field, value):
    raw, _ = _bam_bytes(str(tmp_path))
    payload = bytearray(alnio.bgzf_decompress(raw).tobytes())
    p = _first_record_offset(payload) + 4
    if field == "l_seq":
        struct.pack_into("<i", payload, p + 16, value)
    elif field == "n_cigar":
        struct.pack_into("<H", payload, p + 12, value)
    else:
        payload[p + 8] = value
    with pytest.raises(InputError):
        alnio._read_bam(_recompress(bytes(payload)))


def test_block_size_47_with_huge_l_seq(tmp_path):
    """The advisor's reproducer: block_size 47, l_seq 200 M."""
    raw, _ = _bam_bytes(str(tmp_path), n=2)
    payload = bytearray(alnio.bgzf_decompress(raw).tobytes())
    p = _first_record_offset(payload)
    rec = bytearray(4 + 47)
    struct.pack_into("<I", rec, 0, 47)
    struct.pack_into("<i", rec, 4 + 16, 200_000_000)
    with pytest.raises(InputError):
        alnio._read_bam(_recompress(bytes(payload[:p]) + bytes(rec)))


def test_truncated_and_garbled_bgzf(tmp_path):
    raw, _ = _bam_bytes(str(tmp_path))
    with pytest.raises(InputError):
        alnio._read_bam(raw[:len(raw) // 2])
    bad = bytearray(raw)
    struct.pack_into("<H", bad, 10, 60000)            # XLEN beyond the block
    with pytest.raises(InputError):
        alnio._read_bam(bytes(bad))
    bad = bytearray(raw)
    struct.pack_into("<H", bad, 14, 500)              # subfield length beyond the extra field
    with pytest.raises(InputError):
        alnio._read_bam(bytes(bad))
    bad = bytearray(raw)
    struct.pack_into("<H", bad, 16, 10)               # BSIZE smaller than header + trailer
    with pytest.raises(InputError):
        alnio._read_bam(bytes(bad))


def test_bad_header_lengths(tmp_path):
    raw, _ = _bam_bytes(str(tmp_path))
    payload = bytearray(alnio.bgzf_decompress(raw).tobytes())
    struct.pack_into("<i", payload, 4, 2_000_000_000)
    with pytest.raises(InputError):
        alnio._read_bam(_recompress(bytes(payload)))
