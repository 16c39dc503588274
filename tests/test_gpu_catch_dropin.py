"""-m gpu: the reference's OWN Catch2 suite (20 cases, /root/reference/tests/tests.cpp + tests_statevector.hpp +
tests_densitymatrix.hpp) compiled against the drop-in headers of this repo (tests/catch_dropin/build.sh) and run on the
GPU at 1, 2, 4 and 8 ranks -- the correctness procedure BASELINE north_star names. comm_init() forks the ranks
(DFSA_NP replaces mpirun -np); with fewer GPUs than ranks they share device 0 through the IPC transport.
Comparator: two-sided, 1e-12 relative to max(1, max|ref|) (host/states.hpp agreesWith, -DDFSA_AGREES_TOL=1e-12)."""
import os
import re
import subprocess

import pytest

import product

pytestmark = pytest.mark.gpu

BIN = os.path.join(product.ROOT, "tests", "catch_dropin", "_build", "catch_dropin")
# trials per case: the reference's 5000 at one rank; fewer where 2-8 processes time-share one GPU
TRIALS = {1: 5000, 2: 1500, 4: 600, 8: 300}


@pytest.mark.parametrize("nodes", [1, 2, 4, 8])
def test_reference_catch2_cases_pass_against_the_drop_in_headers(nodes):
    if not os.path.exists(BIN):
        pytest.skip("tests/catch_dropin/_build/catch_dropin is not built (needs /root/reference at build time)")
    env = dict(os.environ)
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    env["DFSA_NP"] = str(nodes)
    env["DFSA_CATCH_TRIALS"] = os.environ.get("DFSA_CATCH_TRIALS_NP%d" % nodes, str(TRIALS[nodes]))
    res = subprocess.run([BIN], env=env, capture_output=True, text=True, timeout=1500)
    tail = (res.stdout + res.stderr)[-3000:]
    assert res.returncode == 0, tail
    m = re.search(r"All tests passed \((\d+) assertions? in (\d+) test cases?\)", res.stdout)
    assert m, tail
    assert int(m.group(2)) == 20, tail
    # 19 cases assert once per trial, expecPauliString twice
    assert int(m.group(1)) == 21 * int(env["DFSA_CATCH_TRIALS"]), tail
