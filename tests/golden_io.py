"""Golden-fixture (tests/golden/*.npz) save/load helpers shared by the generator and the tests."""
import json
import os

import numpy as np

from amplipy_b200.batch import ReadBatch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BATCH_FIELDS = ("pos", "flag", "tlen", "cig_off", "cigar", "seq_off", "seq", "qual_off", "qual")


def save_case(name, batch, meta, arrays):
    """meta: JSON-able dict (params, primers, consensus, variants, insertions ...)."""
    d = {"b_" + f: getattr(batch, f) for f in BATCH_FIELDS}
    d.update(arrays)
    d["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **d)


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    batch = ReadBatch(*[z["b_" + f] for f in BATCH_FIELDS]).validate()
    meta = json.loads(bytes(z["meta_json"]).decode())
    arrays = {k: z[k] for k in z.files if not k.startswith("b_") and k != "meta_json"}
    return batch, meta, arrays


def list_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def ins_from_meta(meta):
    return {(int(p), s): int(c) for p, s, c in meta["insertions"]}
