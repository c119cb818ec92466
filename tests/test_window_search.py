"""CPU: the kernels' sliding-window closed forms (rolling and word-wise dp4a variant) against a literal
python restatement of the reference loops (AmpliPy.py:566-587, 628-649)."""
import ctypes

import numpy as np

import emu_driver


def ref_fwd(qual, width, minq):
    total = 0; true_start = 0; true_end = len(qual); window = min(width, true_end)
    i = true_start
    for offset in range(window - 1):
        total += qual[i + offset]
    while i < true_end:
        if (true_end - window) < i:
            window -= 1
        else:
            total += qual[i + window - 1]
        if (total / window) < minq:
            break
        total -= qual[i]; i += 1
    return true_end - i


def ref_rev(qual, width, minq):
    total = 0; true_start = 0; true_end = len(qual); window = min(width, true_end)
    i = true_end
    for offset in range(1, window):
        total += qual[i - offset]
    while i > true_start:
        if true_start + window > i:
            window -= 1
        else:
            total += qual[i - window]
        if (total / window) < minq:
            break
        total -= qual[i - 1]; i -= 1
    return i


def test_window_closed_forms():
    lib = emu_driver.lib()
    rng = np.random.default_rng(0)
    n_checked = 0
    for trial in range(6000):
        n = int(rng.integers(0, 70))
        mode = trial % 4
        if mode == 0:
            q = rng.integers(0, 42, n)
        elif mode == 1:
            q = rng.choice([37, 25, 11, 2], p=[0.7, 0.2, 0.08, 0.02], size=n)
        elif mode == 2:
            q = np.full(n, 30); q[rng.integers(0, max(n, 1), size=min(n, 3))] = rng.integers(0, 20) if n else 0
        else:
            k = int(rng.integers(0, n + 1)); q = np.concatenate([rng.integers(25, 41, k), rng.integers(0, 18, n - k)])
            if rng.random() < 0.5:
                q = q[::-1]
        q = q.astype(np.uint8)
        minq = int(rng.choice([0, 10, 20, 30]))
        for align in range(4):
            buf = np.full(n + 64, 255 if trial & 1 else 0, np.uint8)
            o = 16 + align + (-buf.ctypes.data) % 4
            buf[o:o + n] = q
            ptr = ctypes.c_void_p(buf.ctypes.data + o)
            for W in (1, 3, 4, 10):
                wf, wr = ref_fwd(q.tolist(), W, minq), ref_rev(q.tolist(), W, minq)
                assert lib.emu_window(ptr, n, W, minq, 0, 1) == wf
                assert lib.emu_window(ptr, n, W, minq, 1, 1) == wr
                if W == 4:
                    assert lib.emu_window(ptr, n, 4, minq, 0, 2) == wf, (q.tolist(), minq, align)
                    assert lib.emu_window(ptr, n, 4, minq, 1, 2) == wr, (q.tolist(), minq, align)
                n_checked += 1
    assert n_checked > 10000
