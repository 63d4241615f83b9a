"""Loader for tests/golden/*.npz (written by oracle/make_golden.py from the reference's twin)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NONE = -(10 ** 9)        # sentinel used in the fixtures for "argument omitted"


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    n = int(z["count"])
    recs = [dict() for _ in range(n)]
    for key in z.files:
        if key == "count":
            continue
        i, k = key.split(":", 1)
        v = z[key]
        if v.ndim == 0:
            v = v.item()
        recs[int(i)][k] = v
    return recs


def opt(v):
    return None if v == NONE else int(v)


def rel_err(a, b):
    """max-norm relative error ||a-b||_inf / ||b||_inf (SURVEY 8c tolerance definition)."""
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.max(np.abs(b)) if b.size else 0.0
    num = np.max(np.abs(a - b)) if b.size else 0.0
    return num / den if den > 0 else num
