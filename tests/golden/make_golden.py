"""Generates tests/golden/golden_{sv,dm}.npz from the REAL reference (oracle/_ref/ref_driver).

Run in the build container (needs /root/reference to have been built by oracle/build_ref.sh):
    python tests/golden/make_golden.py
The reference ships no golden files, so these fixtures are "outputs of the reference itself
run here": seeded random inputs -> unmodified reference API (+ the one-token setBit patch)
at 1, 2, 4 and 8 ranks under the MPI stand-in. Each case stores the input amplitudes, the op parameters
and the reference's output (state, expectation value, or reduced + mutated state for partialTrace).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402
from oracle import refrun  # noqa: E402

TRIALS = 3


def encode_op(op, arrays, prefix):
    enc = []
    for i, a in enumerate(op):
        if isinstance(a, np.ndarray):
            key = "%s_a%d" % (prefix, i)
            arrays[key] = a
            enc.append({"nd": key})
        elif isinstance(a, list) and a and isinstance(a[0], np.ndarray):
            key = "%s_a%d" % (prefix, i)
            arrays[key] = np.stack(a)
            enc.append({"ndlist": key})
        else:
            enc.append(a)
    return enc


def main():
    rng = np.random.default_rng(20261017)
    for kind, nq, names in (("sv", 6, cases.SV_OPS), ("dm", 4, cases.DM_OPS)):
        nbits = nq if kind == "sv" else 2 * nq
        arrays, meta = {}, []
        cid = 0
        for name in names:
            for nodes in (1, 2, 4, 8):
                for _ in range(TRIALS):
                    op = cases.make_op(rng, name, nq, nodes.bit_length() - 1)
                    amps = cases.random_state(rng, nbits)
                    ref = refrun.run(kind, nq, [op], num_nodes=nodes, init_amps=amps)
                    p = "c%d" % cid
                    arrays[p + "_in"] = amps
                    entry = {"id": cid, "kind": kind, "nq": nq, "nodes": nodes, "op": encode_op(op, arrays, p)}
                    if name == "dm_expecPauliString":
                        arrays[p + "_val"] = np.array(ref["values"][:1])
                    else:
                        arrays[p + "_out"] = ref["amps"]
                    if name == "dm_partialTrace":
                        arrays[p + "_mut"] = ref["mutated"][0]
                    meta.append(entry)
                    cid += 1
        arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        path = os.path.join(HERE, "golden_%s.npz" % kind)
        np.savez_compressed(path, **arrays)
        print("wrote %s: %d cases, %.1f KiB" % (path, cid, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
