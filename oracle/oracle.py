"""ctypes wrapper around ``oracle/amplipy_oracle.c`` -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product package never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libamplipy_oracle.so")
_lib = None

F_TRIM_START, F_TRIM_END, F_TRIM_QUAL, F_KEEP, F_SKIPPED, F_ERROR = 1, 2, 4, 8, 16, 32
FIXED_SYMS = "ACGTN-"


def build(force=False):
    src = os.path.join(_HERE, "amplipy_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


def use_native_build():
    """bench.py's CPU legs: build the same source with -O3 -march=native ON THE MACHINE THAT RUNS IT (the portable
    -O2 build is what travels with the repository) and load that one.  Call before the first use of the library."""
    global _SO, _lib
    src = os.path.join(_HERE, "amplipy_oracle.c")
    so = os.path.join(_HERE, "_build", "libamplipy_oracle_native.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    cc = "/usr/bin/gcc" if os.path.isfile("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-O3", "-march=native", "-fPIC", "-fopenmp", "-shared", "-o", so, src])
    _SO, _lib = so, None
    return so


def set_num_threads(n):
    lib().oracle_set_num_threads(ctypes.c_int(int(n)))


def _p(a, ct=None):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.c_void_p)


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_pileup_batch.restype = ctypes.c_void_p
        _lib.oracle_ins_count.restype = ctypes.c_int64
        _lib.oracle_ins_chars.restype = ctypes.c_int64
        _lib.oracle_call.restype = ctypes.c_int64
        _lib.oracle_ins_count.argtypes = [ctypes.c_void_p]
        _lib.oracle_ins_chars.argtypes = [ctypes.c_void_p]
        _lib.oracle_pileup_free.argtypes = [ctypes.c_void_p]
        _lib.oracle_ins_export.argtypes = [ctypes.c_void_p] * 5
    return _lib


def num_threads():
    return int(lib().oracle_num_threads())


def find_overlapping_primers(L, primers, offset):
    """primers: sorted list of (start, end).  Returns int32 tables with -1 for None."""
    st = np.array([p[0] for p in primers], np.int32)
    en = np.array([p[1] for p in primers], np.int32)
    mn = np.empty(L, np.int32)
    mx = np.empty(L, np.int32)
    lib().oracle_find_overlapping_primers(ctypes.c_int(L), ctypes.c_int(len(primers)), _p(st), _p(en),
                                          ctypes.c_int(offset), _p(mn), _p(mx))
    return mn, mx


def trim_batch(b, L, min_start, max_end, max_primer_len, min_quality=20, window=4, min_length=30,
               include_no_primer=False):
    n = b.n
    out_pos = np.empty(n, np.int32)
    out_ncig = np.empty(n, np.int32)
    out_cigar = np.zeros(int(b.cig_off[-1]) + 3 * n, np.uint32)
    out_flags = np.empty(n, np.uint8)
    lib().oracle_trim_batch(ctypes.c_int64(n), _p(b.pos), _p(b.flag), _p(b.tlen), _p(b.cig_off), _p(b.cigar),
                            _p(b.qual_off), _p(b.qual), ctypes.c_int(L), _p(min_start), _p(max_end),
                            ctypes.c_int(max_primer_len), ctypes.c_int(min_quality), ctypes.c_int(window),
                            ctypes.c_int(min_length), ctypes.c_int(1 if include_no_primer else 0),
                            _p(out_pos), _p(out_ncig), _p(out_cigar), _p(out_flags))
    return {"pos": out_pos, "ncig": out_ncig, "cigar": out_cigar, "flags": out_flags}


def trimmed_cigartuples(b, t, i):
    a = int(b.cig_off[i]) + 3 * i
    return [(int(c & 15), int(c >> 4)) for c in t["cigar"][a:a + int(t["ncig"][i])]]


def pileup_batch(b, L, min_quality=20, trimmed=None):
    """Returns (counts int64[6, L], {(pos, str): count}, n_errors).  ``trimmed`` = trim_batch() output
    to pile up the trimmed alignments (the `aio` data flow, AmpliPy.py:907-915)."""
    counts = np.zeros((6, L), np.int64)
    nerr = ctypes.c_int64(0)
    if trimmed is None:
        pos, cigar, ncig, stride3, skip = b.pos, b.cigar, None, 0, None
    else:
        pos, cigar, ncig, stride3, skip = trimmed["pos"], trimmed["cigar"], trimmed["ncig"], 1, trimmed["flags"]
    h = lib().oracle_pileup_batch(ctypes.c_int64(b.n), _p(pos), _p(b.flag), _p(b.cig_off), _p(cigar), _p(ncig),
                                  ctypes.c_int(stride3), _p(skip), _p(b.seq_off), _p(b.seq), _p(b.qual_off),
                                  _p(b.qual), ctypes.c_int(L), ctypes.c_int(min_quality), _p(counts),
                                  ctypes.byref(nerr))
    h = ctypes.c_void_p(h)
    k = int(lib().oracle_ins_count(h))
    nch = int(lib().oracle_ins_chars(h))
    ipos = np.empty(k, np.int32)
    icnt = np.empty(k, np.int64)
    ioff = np.empty(k + 1, np.int64)
    chars = np.empty(max(nch, 1), np.uint8)
    lib().oracle_ins_export(h, _p(ipos), _p(icnt), _p(ioff), _p(chars))
    lib().oracle_pileup_free(h)
    raw = chars.tobytes()
    ins = {}
    for j in range(k):
        ins[(int(ipos[j]), raw[int(ioff[j]):int(ioff[j + 1])].decode())] = int(icnt[j])
    return counts, ins, int(nerr.value)


def call(counts, ins, ref_seq, run_consensus=True, min_depth_consensus=10, min_freq_consensus=0.0,
         run_variants=True, min_depth_variants=1, min_freq_variants=0.03):
    """Restates the per-position loop AmpliPy.py:921-952.  ``ins`` = {(pos, str): count}.
    Returns a dict of python structures comparable with the reference's own values."""
    L = counts.shape[1]
    items = sorted(ins.items(), key=lambda kv: kv[0][0])
    K = len(items)
    ipos = np.array([k[0] for k, _ in items], np.int32)
    icnt = np.array([v for _, v in items], np.int64)
    strs = [k[1].encode() for k, _ in items]
    ioff = np.zeros(K + 1, np.int64)
    if K:
        np.cumsum([len(s) for s in strs], out=ioff[1:])
    chars = np.frombuffer(b"".join(strs) + b"\0", np.uint8).copy()
    cap = int(np.count_nonzero(counts)) + K + 1
    depth = np.empty(L, np.int64)
    al_off = np.empty(L + 1, np.int64)
    al_count = np.empty(cap, np.int64)
    al_freq = np.empty(cap, np.float64)
    al_sym = np.empty(cap, np.int32)
    al_is_alt = np.empty(cap, np.uint8)
    cons_sym = np.empty(L, np.int32)
    var_emit = np.empty(L, np.uint8)
    var_gt_ref = np.empty(L, np.uint8)
    var_ref_count = np.empty(L, np.int64)
    var_ref_freq = np.empty(L, np.float64)
    refb = np.frombuffer(ref_seq.encode(), np.uint8).copy()
    cnt = np.ascontiguousarray(counts, np.int64)
    lib().oracle_call(ctypes.c_int(L), _p(cnt), ctypes.c_int64(K), _p(ipos), _p(icnt), _p(ioff), _p(chars), _p(refb),
                      ctypes.c_int(int(run_consensus)), ctypes.c_int64(min_depth_consensus),
                      ctypes.c_double(min_freq_consensus), ctypes.c_int(int(run_variants)),
                      ctypes.c_int64(min_depth_variants), ctypes.c_double(min_freq_variants),
                      _p(depth), _p(al_off), _p(al_count), _p(al_freq), _p(al_sym), _p(al_is_alt), _p(cons_sym),
                      _p(var_emit), _p(var_gt_ref), _p(var_ref_count), _p(var_ref_freq))

    def sym(i):
        return FIXED_SYMS[i] if i < 6 else strs[i - 6].decode()
    return {"depth": depth, "al_off": al_off, "al_count": al_count, "al_freq": al_freq, "al_sym": al_sym,
            "al_is_alt": al_is_alt, "cons_sym": cons_sym, "var_emit": var_emit, "var_gt_ref": var_gt_ref,
            "var_ref_count": var_ref_count, "var_ref_freq": var_ref_freq, "sym": sym, "ins_strs": strs}


def consensus_string(res, unknown_symbol="N"):
    L = res["depth"].shape[0]
    return "".join(unknown_symbol if res["cons_sym"][p] < 0 else res["sym"](int(res["cons_sym"][p])) for p in range(L))


def variant_records(res, ref_seq):
    """[(pos0, ref, [alts], DP, REF_DP, [ALT_DP], REF_FREQ, [ALT_FREQ], gt_tuple)] as AmpliPy.py:941-951."""
    out = []
    for p in np.flatnonzero(res["var_emit"]):
        a, b = int(res["al_off"][p]), int(res["al_off"][p + 1])
        alts = [i for i in range(a, b) if res["al_is_alt"][i]]
        n = len(alts)
        gt = tuple(range(n + 1)) if res["var_gt_ref"][p] else tuple(range(1, n + 1))
        out.append((int(p), ref_seq[p], [res["sym"](int(res["al_sym"][i])) for i in alts], int(res["depth"][p]),
                    int(res["var_ref_count"][p]), [int(res["al_count"][i]) for i in alts],
                    float(res["var_ref_freq"][p]), [float(res["al_freq"][i]) for i in alts], gt))
    return out
