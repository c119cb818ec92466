"""Small end-to-end case for compute-sanitizer runs (memcheck / racecheck) on the GPU box:
    compute-sanitizer --tool memcheck python tests/sanitizer_case.py
Covers both kernel variants (short-read and indel-rich), the calling kernels and the insertion merge."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_io  # noqa: E402
import parity  # noqa: E402
from amplipy_b200.engine import Engine  # noqa: E402

for name in ("cfg2_illumina", "cfg4_ont", "fuzz1", "quirks"):
    eng = parity.check_case_aio(Engine, name)
    ins = eng.insertions()
    eng.merge_insertions(ins.sample, ins.pos, ins.count, ins.str_off, ins.chars)
    assert eng.error_flags() == 0
    eng.close()
print("sanitizer case ok")
