"""Test-side access to the product (the native CUDA path through the host API)."""
import importlib
import os
import pickle
import subprocess
import sys
import tempfile
import uuid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "distributed-full-state-algorithms_b200"


def pkg():
    return importlib.import_module(PKG_NAME)


def gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=60)
        return sum(1 for line in out.stdout.splitlines() if line.startswith("GPU "))
    except Exception:
        return 0


def run_cases_multirank(cases, nodes, timeout=600, extra_env=None):
    """Run `cases` (dicts with kind, nq, op, amps) at `nodes` ranks, one process per rank.
    With >= nodes GPUs the ranks use NCCL (one GPU each), otherwise the IPC transport on GPU 0.
    Returns the list of per-case results as produced by tests/mp_worker.py on rank 0."""
    with tempfile.TemporaryDirectory(prefix="dfsa_mp_") as tmp:
        job = os.path.join(tmp, "job.pkl")
        out = os.path.join(tmp, "out.pkl")
        with open(job, "wb") as f:
            pickle.dump(cases, f)
        job_id = uuid.uuid4().hex[:16]
        procs = []
        for r in range(nodes):
            env = dict(os.environ)
            env.update({"RANK": str(r), "WORLD_SIZE": str(nodes), "LOCAL_RANK": str(r), "DFSA_JOB_ID": job_id})
            if extra_env:
                env.update(extra_env)
            procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mp_worker.py"), job, out],
                                          env=env, stdout=open(os.path.join(tmp, "rank%d.log" % r), "w+"), stderr=subprocess.STDOUT, text=True))
        # poll: the first rank that dies takes the job down (its partners would otherwise sit in a barrier)
        import time
        deadline = time.time() + timeout
        failed = None
        while True:
            codes = [p.poll() for p in procs]
            if any(c is not None and c != 0 for c in codes):
                failed = "rank %d exited with %d" % (next(i for i, c in enumerate(codes) if c not in (None, 0)), next(c for c in codes if c not in (None, 0)))
            elif all(c == 0 for c in codes):
                break
            elif time.time() > deadline:
                failed = "timeout after %ds" % timeout
            if failed:
                time.sleep(1.0)
                for q in procs:
                    if q.poll() is None:
                        q.kill()
                break
            time.sleep(0.05)
        logs = []
        for r, p in enumerate(procs):
            try:
                p.wait(timeout=30)
            except subprocess.TimeoutExpired:
                p.kill()
            with open(os.path.join(tmp, "rank%d.log" % r)) as f:
                logs.append("[rank %d] " % r + f.read()[-4000:])
        if failed:
            raise RuntimeError("multi-rank run failed (%s):\n" % failed + "\n---\n".join(logs))
        with open(out, "rb") as f:
            return pickle.load(f)
