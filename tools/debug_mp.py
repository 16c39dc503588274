"""Debug helper: run the first few multi-rank golden cases with tracing and a short timeout."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_io, product
nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
limit = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
todo = [c for c in golden_io.load("sv") + golden_io.load("dm") if c["nodes"] == nodes][:limit]
jobs = [dict(kind=c["kind"], nq=c["nq"], op=c["op"], amps=c["amps"]) for c in todo]
try:
    res = product.run_cases_multirank(jobs, nodes, timeout=90, extra_env={"DFSA_TRACE": "1", "DFSA_COMM_TIMEOUT_S": "20"})
    print("ok", len(res), res[0]["transport"])
except RuntimeError as e:
    print(str(e)[-6000:])
