"""Loading of the committed golden fixtures (tests/golden/*.npz, made by oracle/gen_golden.py)."""
import glob
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLD, "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(str(d["meta"]))
    d["encoded"] = d["encoded"].tobytes()
    if "mv_txt" in d:
        d["mv_txt"] = str(d["mv_txt"])
    return d


def split_container(data: bytes):
    """Parse the container (encoder.py:104-121) into [(mode, pred_bytes, coef_bytes)]."""
    out, o = [], 0
    while o < len(data):
        mode = data[o]
        pl = int.from_bytes(data[o + 1:o + 3], "big")
        pred = data[o + 3:o + 3 + pl]
        o += 3 + pl
        cl = int.from_bytes(data[o:o + 3], "big")
        coef = data[o + 3:o + 3 + cl]
        o += 3 + cl
        out.append((mode, pred, coef))
    return out
