"""One rank of a multi-process product run (spawned by tests/product.py: RANK/WORLD_SIZE/DFSA_JOB_ID in env).
Replays each case on the CUDA path and lets rank 0 write the gathered results."""
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import cases as cases_mod  # noqa: E402
import product  # noqa: E402


def main():
    job, out_path = sys.argv[1], sys.argv[2]
    with open(job, "rb") as f:
        todo = pickle.load(f)
    dfsa = product.pkg()
    dfsa.comm_init()
    results = []
    for ci, case in enumerate(todo):
        if os.environ.get("DFSA_TRACE"):
            print("[worker %d] case %d: %s" % (dfsa.comm_rank(), ci, (case["ops"][0][0] if "ops" in case else case["op"][0])), flush=True)
        st = dfsa.DeviceState(case["kind"], case["nq"])
        if case.get("amps") is not None:
            st.set_amps(case["amps"])
        else:
            st.init_hash(case["seed"])
        res = {"values": []}
        ops = case["ops"] if "ops" in case else [case["op"]]
        cur = st
        for op in ops:
            r = cases_mod.apply(cur, op)
            if op[0] == "dm_expecPauliString":
                res["values"].append(r)
            elif op[0] == "dm_partialTrace":
                res["mutated"] = cur.get_amps()
                cur = r
        res["layout_before_readback"] = cur.layout()
        res["amps"] = cur.get_amps()
        res["layout_after_readback"] = cur.layout()
        res["transport"] = dfsa.device_lib().dfsa_comm_transport().decode()
        results.append(res)
        if cur is not st:
            cur.close()
        st.close()
    if dfsa.comm_rank() == 0:
        with open(out_path, "wb") as f:
            pickle.dump(results, f)
    dfsa.comm_end()


if __name__ == "__main__":
    main()
