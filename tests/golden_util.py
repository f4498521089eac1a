"""Helpers shared by the parity tests: load a golden fixture written by oracle/gen_golden.py."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES_ATT = ["s2vtatt_tiny", "s2vtatt_mid"]
CASES_S2VT = ["s2vt_tiny", "s2vt_mid", "s2vt_sched"]
CASES_RAT = ["rationale_att_tiny", "rationale_att_mid", "rationale_s2vt_tiny"]


def load(tag, dtype=np.float64):
    z = np.load(os.path.join(GOLDEN, tag + ".npz"))
    d = {k: z[k] for k in z.files}
    params = {k[6:]: d[k].astype(dtype) for k in d if k.startswith("param.")}
    grads = {k[5:]: d[k] for k in d if k.startswith("grad.")}
    return d, params, grads


def relerr(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
