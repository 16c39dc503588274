"""Loader for the committed golden fixtures (tests/golden/golden_{sv,dm}.npz, made by make_golden.py)."""
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def load(kind):
    if kind not in _cache:
        z = np.load(os.path.join(_HERE, "golden", "golden_%s.npz" % kind))
        meta = json.loads(bytes(z["meta"]).decode())
        out = []
        for m in meta:
            p = "c%d" % m["id"]
            op = []
            for a in m["op"]:
                if isinstance(a, dict) and "nd" in a:
                    op.append(z[a["nd"]])
                elif isinstance(a, dict) and "ndlist" in a:
                    op.append(list(z[a["ndlist"]]))
                else:
                    op.append(a)
            case = dict(id=m["id"], kind=m["kind"], nq=m["nq"], nodes=m["nodes"], op=tuple(op), amps=z[p + "_in"])
            for suffix in ("out", "val", "mut"):
                if p + "_" + suffix in z:
                    case[suffix] = z[p + "_" + suffix]
            out.append(case)
        _cache[kind] = out
    return _cache[kind]


def case_id(case):
    return "%s-%s-np%d-#%d" % (case["kind"], case["op"][0], case["nodes"], case["id"])
