"""CPU test of the drop-in host layer's decisions at 1...16 ranks (no GPU): random circuits through host/*.hpp -- local vs.
exchange, partner ranks, relocation plans, the lazy qubit layout, swap-in of rank-bit qubits and which qubit it evicts, the
deferred gate queue, layout restores, partialTrace's reordering -- running on tests/hostsim/dfsa_hostsim.cpp, a CPU stand-in
that implements what include/dfsa_b200.h documents for each C-ABI entry (one forked process per rank, pairwise steps
synchronised pairwise), compared with a dense ground truth that knows nothing about ranks (tests/hostsim/fuzz.cpp).

The stand-in is test infrastructure: nothing in the package or bench.py links or loads it (checked below), and the GPU
suite is what proves the kernels. What this adds is breadth the GPU box has no time for: thousands of random circuits per rank
count, in every layout / fusion mode, where a wrong host decision shows up as a mismatch or as "rank r waits for rank p".
"""
import os
import re
import signal
import subprocess

import pytest

import product

HERE = os.path.join(product.ROOT, "tests", "hostsim")
BIN = os.path.join(HERE, "_build", "fuzz")


@pytest.fixture(scope="module")
def fuzz_binary():
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.run(["bash", os.path.join(HERE, "build.sh")], check=True, env=env, stdout=subprocess.DEVNULL)
    return BIN


def run_fuzz(binary, nodes, trials, seed, lazy, fuse):
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    env.update(DFSA_NP=str(nodes), DFSA_LAZY_LAYOUT=str(lazy), DFSA_FUSE_GATES=str(fuse))
    proc = subprocess.Popen([binary, str(trials), str(seed)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = proc.communicate(timeout=240)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)            # the rank processes it forked share its process group
        out, err = proc.communicate()
        pytest.fail("hostsim fuzz hung at %d ranks:\n%s" % (nodes, err[-3000:]))
    return proc.returncode, out, err


@pytest.mark.parametrize("lazy,fuse", [(1, 1), (0, 1), (1, 0), (0, 0)])
@pytest.mark.parametrize("nodes", [1, 2, 4, 8, 16])
def test_random_circuits_through_the_host_layer_match_dense_truth(fuzz_binary, nodes, lazy, fuse):
    trials = 600 if (lazy, fuse) == (1, 1) else 250
    rc, out, err = run_fuzz(fuzz_binary, nodes, trials, seed=20 + nodes, lazy=lazy, fuse=fuse)
    assert rc == 0, (out + err)[-3000:]
    m = re.search(r"P=(\d+) lazy_layout=(\d) gate_fusion=(\d) trials=(\d+) comparisons=(\d+) failures=(\d+)", out)
    assert m, out[-2000:]
    assert (int(m.group(1)), int(m.group(2)), int(m.group(3)), int(m.group(4))) == (nodes, lazy, fuse, trials)
    assert int(m.group(5)) >= trials * nodes and int(m.group(6)) == 0          # every rank compared at least once per trial
    # the paths this test is about were taken (calls summed over ranks)
    calls = {k: int(v) for k, v in re.findall(r" (\w+)=(\d+)", out.split("calls, all ranks:")[1])}
    assert calls["k_manyTarg"] > 0 and calls["k_pauli"] > 0 and calls["k_partialTrace"] > 0 and calls["k_krausMap"] > 0
    assert (calls["k_gateSequence"] > 0) == bool(fuse) and (calls["k_ctrlOneTarg"] > 0) == (not fuse)
    if nodes > 1:
        assert calls["xk_relocate"] > 0 and calls["xk_swapSuffixPrefix"] > 0 and calls["xk_exchangePauliCombine"] > 0
        if not (lazy and fuse):        # with both on, one-target gates on rank-bit qubits swap the qubit in instead of exchanging shards
            assert calls["xk_exchangeCombine"] > 0 and calls["xk_ctrlPrefixTarg"] > 0
        else:
            assert calls["xk_exchangeCombine"] == 0 and calls["xk_ctrlPrefixTarg"] == 0
    if nodes > 2:
        assert calls["x_exchange"] > 0                                         # rank-bit <-> rank-bit swaps (swapGate, or a layout restore)


@pytest.mark.parametrize("nodes", [1, 2, 4, 8])
def test_host_c_api_on_the_stand_in(fuzz_binary, nodes):
    """The extern "C" face of the host layer (what api.py / bench.py call through ctypes) with a live StateVector: pending gates,
    the flush plan of a state (decoded by api.py's decoder, equal to the stateless planner's), layouts before / after the flush,
    amplitudes against numpy -- tests/hostsim/capi_on_standin.py, every rank running the script."""
    import sys
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "DFSA_LAZY_LAYOUT", "DFSA_FUSE_GATES"):
        env.pop(k, None)
    env["DFSA_NP"] = str(nodes)
    proc = subprocess.Popen([sys.executable, os.path.join(HERE, "capi_on_standin.py")], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = proc.communicate(timeout=240)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        out, err = proc.communicate()
        pytest.fail("capi_on_standin.py hung at %d ranks:\n%s" % (nodes, err[-3000:]))
    assert proc.returncode == 0 and "capi on stand-in: ok P=%d" % nodes in out, (out + err)[-3000:]


@pytest.mark.parametrize("nodes", [1, 2, 8])
def test_bench_product_arm_dry_run_on_the_stand_in(fuzz_binary, nodes):
    """bench.py's own control flow (warm-up, self-check, timed steps taking the flush plan of every step, fused-pass accounting,
    per-gate mode, exchange summary, end-to-end leg, the JSON line) executed on the CPU at a toy size -- tests/hostsim/
    bench_dry_run.py. The numbers are meaningless and go nowhere; the line must exist, carry the contract's keys, pass its own
    full-state self-check (sweep, then inverse, against the regenerated state) and account for the relocations it planned."""
    import json
    import sys
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "DFSA_LAZY_LAYOUT", "DFSA_FUSE_GATES"):
        env.pop(k, None)
    env["DFSA_NP"] = str(nodes)
    proc = subprocess.Popen([sys.executable, os.path.join(HERE, "bench_dry_run.py"), "12", "configs"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = proc.communicate(timeout=240)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        out, err = proc.communicate()
        pytest.fail("bench_dry_run.py hung at %d ranks:\n%s" % (nodes, err[-3000:]))
    assert proc.returncode == 0, (out + err)[-3000:]
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out[-2000:]
    line = json.loads(lines[0])
    k = nodes.bit_length() - 1
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "roofline", "step_roofline", "per_gate_mode", "e2e", "gpu_launches", "clocks", "selfcheck"):
        assert key in line, key
    assert line["n_gpus"] == nodes and line["config"]["qubits"] == 12 + k and line["config"]["gates_per_step"] == 2 * (12 + k)
    assert line["selfcheck"]["ok"] and line["gate_fusion"] is True
    roof = line["roofline"]
    assert "accounting_error" not in roof and len(roof["passes_by_step"]) == line["steps"] and all(p >= 1 for p in roof["passes_by_step"])
    for step in roof["relocation_pairs_by_step"]:
        assert (len(step) >= 1) == (nodes > 1) and all(1 <= m <= max(k, 1) for m in step)
    assert 0 < roof["frac"] and 0 < line["step_roofline"]["frac"] and line["e2e"]["h2d_bytes_per_step"] == 16 * (1 << (12 + k))
    assert (line["exchange_gates"] is not None) == (nodes > 1)
    # BASELINE configs 3-5 at toy sizes (same op mix): bench.run_config's cost model, restore plans and per-op tables at this rank count
    configs = line["configs"]
    assert "config3_circuit" in configs and any(name.startswith("config4_") for name in configs) and any(name.startswith("config5_") for name in configs)
    for name, c in configs.items():
        assert c["roofline_frac"] > 0 and c["per_op"], name
        labels = " ".join(c["per_op"])
        if nodes > 1 and name.startswith("config3"):
            assert "[exchange]" in labels
        if nodes > 1 and name.startswith("config5"):
            assert "(relocating)" in labels and "(local)" in labels


@pytest.mark.parametrize("nodes", [1, 2, 4, 8, 16])
def test_reference_catch2_cases_pass_on_the_host_layer_over_the_stand_in(fuzz_binary, nodes):
    """The reference's OWN 20 Catch2 cases (compiled against host/*.hpp by tests/catch_dropin/build.sh) linked against the CPU
    stand-in instead of libdfsa_b200.so: every API function's host-side dispatch -- including the noise channels' prefix / suffix
    branches, expecPauliString's reduction and partialTrace's relocation -- at up to 16 ranks, with this repo's two-sided 1e-12
    comparator against the suite's dense Kraus-operator ground truth. (The kernels' own parity is the GPU run of the same suite,
    tests/test_gpu_catch_dropin.py.)"""
    binary = os.path.join(HERE, "_build", "catch_on_standin")
    if not os.path.exists(binary):
        pytest.skip("tests/catch_dropin/_build objects are not there (they need /root/reference at build time)")
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "DFSA_LAZY_LAYOUT", "DFSA_FUSE_GATES"):
        env.pop(k, None)
    trials = 400
    env.update(DFSA_NP=str(nodes), DFSA_CATCH_TRIALS=str(trials))
    proc = subprocess.Popen([binary], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = proc.communicate(timeout=280)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        out, err = proc.communicate()
        pytest.fail("catch_on_standin hung at %d ranks:\n%s" % (nodes, (out + err)[-3000:]))
    assert proc.returncode == 0, (out + err)[-3000:]
    m = re.search(r"All tests passed \((\d+) assertions? in (\d+) test cases?\)", out)
    assert m and int(m.group(2)) == 20 and int(m.group(1)) == 21 * trials, out[-2000:]


@pytest.mark.parametrize("nodes", [4, 16])
def test_random_circuits_with_rank_bit_qubits_brought_in_one_pair_at_a_time(fuzz_binary, nodes):
    """DFSA_GROUP_SWAPIN=0: the gate queue's launch plan with one (suffix, rank) pair per relocation step (the A/B switch)."""
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "DFSA_LAZY_LAYOUT", "DFSA_FUSE_GATES"):
        env.pop(k, None)
    env.update(DFSA_NP=str(nodes), DFSA_GROUP_SWAPIN="0")
    proc = subprocess.Popen([fuzz_binary, "250", "9"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = proc.communicate(timeout=240)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        out, err = proc.communicate()
        pytest.fail("hostsim fuzz hung at %d ranks:\n%s" % (nodes, err[-3000:]))
    assert proc.returncode == 0 and "failures=0" in out, (out + err)[-3000:]


@pytest.mark.parametrize("nodes", [2, 4, 8, 16])
def test_bench_nvlink_cost_model_equals_what_the_stand_in_pulls_over_the_link(fuzz_binary, nodes):
    """Every roofline fraction of the bench line divides an algorithmic cost by a time; the NVLink part of that cost (bench.op_cost,
    restore_cost, relocation_cost, the fused-step accounting) is held against the stand-in's own count of the amplitudes each rank
    pulls from other ranks, op by op, for the op mix of BASELINE configs 3-5, layout restores and the sweep in both modes --
    tests/hostsim/cost_model_check.py."""
    import sys
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "DFSA_LAZY_LAYOUT", "DFSA_FUSE_GATES", "DFSA_GROUP_SWAPIN"):
        env.pop(k, None)
    env["DFSA_NP"] = str(nodes)
    proc = subprocess.Popen([sys.executable, os.path.join(HERE, "cost_model_check.py")], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = proc.communicate(timeout=240)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        out, err = proc.communicate()
        pytest.fail("cost_model_check.py hung at %d ranks:\n%s" % (nodes, err[-3000:]))
    assert proc.returncode == 0, (out + err)[-3000:]
    m = re.search(r"cost model check: ok at (\d+) rank\(s\), (\d+) ops compared, (\d+) of them with NVLink traffic", out)
    assert m and int(m.group(1)) == nodes and int(m.group(2)) >= 100 and int(m.group(3)) >= 30, out[-2000:]


def test_the_stand_in_is_not_part_of_the_product():
    """Nothing under the package, include/ or bench.py names the stand-in, and the product's host library links libdfsa_b200."""
    for base, _, files in os.walk(os.path.join(product.ROOT, "distributed-full-state-algorithms_b200")):
        for f in files:
            if f.endswith((".py", ".hpp", ".cpp", ".cu", ".cuh", ".h")):
                with open(os.path.join(base, f), errors="ignore") as fh:
                    assert "hostsim" not in fh.read(), os.path.join(base, f)
    for f in ("bench.py", os.path.join("include", "dfsa_b200.h")):
        with open(os.path.join(product.ROOT, f)) as fh:
            assert "hostsim" not in fh.read(), f
    host_lib = os.path.join(product.ROOT, "distributed-full-state-algorithms_b200", "libdfsa_host.so")
    if os.path.exists(host_lib):
        needed = subprocess.run(["readelf", "-d", host_lib], capture_output=True, text=True).stdout
        assert "libdfsa_b200.so" in needed and "hostsim" not in needed
