"""Runs an op list through the REAL reference build (oracle/_ref/ref_driver, made by build_ref.sh).
CHECKER / CPU-BASELINE ONLY.  The binary is the unmodified reference API + the one-token setBit patch
 + the fork/shared-memory MPI stand-in (SHIM_NP replaces mpirun -np).
"""
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def driver_path():
    """Pick the AVX-512 build when the host CPU has it, else the x86-64-v3 one; None if not built."""
    v3 = os.path.join(_HERE, "_ref", "ref_driver")
    v4 = os.path.join(_HERE, "_ref", "ref_driver_v4")
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    need = ("avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl")
    if os.path.exists(v4) and all(x in flags.split() for x in need):
        return v4
    return v3 if os.path.exists(v3) else None


def available():
    return driver_path() is not None


def _hx(x):
    return float(x).hex()


def _cplx(z):
    z = complex(z)
    return "%s %s" % (_hx(z.real), _hx(z.imag))


def _matrix(m):
    m = np.asarray(m, dtype=np.complex128)
    return " ".join(_cplx(z) for z in m.reshape(-1))


def _ints(xs):
    return " ".join(str(int(x)) for x in xs)


def op_line(op, mutated_path="-"):
    """Serialise one op tuple into the ref_driver grammar (oracle/ref_driver.cpp header)."""
    name = op[0]
    if name in ("sv_oneTargGate",):
        return "%s %d %s" % (name, op[1], _matrix(op[2]))
    if name == "sv_manyCtrlOneTargGate":
        return "%s %d %s %d %s" % (name, len(op[1]), _ints(op[1]), op[2], _matrix(op[3]))
    if name in ("sv_swapGate", "dm_swapGate"):
        return "%s %d %d" % (name, op[1], op[2])
    if name in ("sv_manyTargGate", "dm_manyTargGate"):
        return "%s %d %s %s" % (name, len(op[1]), _ints(op[1]), _matrix(op[2]))
    if name in ("sv_pauliTensor", "dm_pauliTensor"):
        return "%s %d %s %s" % (name, len(op[1]), _ints(op[1]), _ints(op[2]))
    if name in ("sv_pauliGadget", "dm_pauliGadget"):
        return "%s %d %s %s %s" % (name, len(op[1]), _ints(op[1]), _ints(op[2]), _hx(op[3]))
    if name in ("sv_phaseGadget", "dm_phaseGadget"):
        return "%s %d %s %s" % (name, len(op[1]), _ints(op[1]), _hx(op[2]))
    if name == "dm_krausMap":
        return "%s %d %s %d %s" % (name, len(op[1]), _ints(op[1]), len(op[2]), " ".join(_matrix(k) for k in op[2]))
    if name in ("dm_oneQubitDephasing", "dm_oneQubitDepolarising", "dm_damping"):
        return "%s %d %s" % (name, op[1], _hx(op[2]))
    if name in ("dm_twoQubitDephasing", "dm_twoQubitDepolarising"):
        return "%s %d %d %s" % (name, op[1], op[2], _hx(op[3]))
    if name == "dm_expecPauliString":
        coeffs = np.asarray(op[1], dtype=np.float64).reshape(-1)
        return "%s %d %s %s" % (name, coeffs.size, " ".join(_hx(c) for c in coeffs), _ints(np.asarray(op[2]).reshape(-1)))
    if name == "dm_partialTrace":
        return "%s %d %s %s" % (name, len(op[1]), _ints(op[1]), mutated_path)
    raise ValueError("unknown op %r" % (name,))


def run(kind, num_qubits, ops, num_nodes=1, init_amps=None, init_seed=None, threads=None, want_state=True, timed=False, timeout=3600):
    """Execute `ops` on the reference at `num_nodes` ranks.

    Returns dict(amps=final global amplitude array or None, values=[expecPauliString results],
    mutated=[the mutated input of each partialTrace], seconds=wall time of the op list when timed).
    """
    drv = driver_path()
    if drv is None:
        raise RuntimeError("oracle/_ref/ref_driver is not built (run oracle/build_ref.sh where /root/reference exists)")
    with tempfile.TemporaryDirectory(prefix="dfsa_ref_") as tmp:
        lines = ["state %s %d" % (kind, num_qubits)]
        if init_amps is not None:
            np.ascontiguousarray(init_amps, dtype=np.complex128).tofile(os.path.join(tmp, "init.bin"))
            lines.append("init file %s" % os.path.join(tmp, "init.bin"))
        elif init_seed is not None:
            lines.append("init hash %d" % init_seed)
        if timed:
            lines.append("tic")
        n_mut = 0
        for op in ops:
            mp = "-"
            if op[0] == "dm_partialTrace" and not timed:
                mp = os.path.join(tmp, "mut%d.bin" % n_mut)
                n_mut += 1
            lines.append(op_line(op, mp))
        if timed:
            lines.append("toc ops")
        if want_state:
            lines.append("dump %s" % os.path.join(tmp, "out.bin"))
        lines.append("dumpvals %s" % os.path.join(tmp, "vals.bin"))
        script = os.path.join(tmp, "script.txt")
        with open(script, "w") as f:
            f.write("\n".join(lines) + "\n")
        env = dict(os.environ)
        env["SHIM_NP"] = str(num_nodes)
        ncpu = os.cpu_count() or 1
        env["OMP_NUM_THREADS"] = str(threads if threads else max(1, ncpu // num_nodes))
        # compile.sh:11-13 of the reference
        env.setdefault("OMP_WAIT_POLICY", "active")
        env.setdefault("OMP_DYNAMIC", "false")
        env.setdefault("OMP_PROC_BIND", "true")   # the shim gives each rank its own core slice
        res = subprocess.run([drv, script], env=env, capture_output=True, text=True, timeout=timeout)
        if res.returncode != 0:
            raise RuntimeError("ref_driver failed (%d): %s" % (res.returncode, res.stderr[-2000:]))
        out = {"amps": None, "values": [], "mutated": [], "seconds": None, "threads": int(env["OMP_NUM_THREADS"])}
        if want_state:
            out["amps"] = np.fromfile(os.path.join(tmp, "out.bin"), dtype=np.complex128)
        out["values"] = list(np.fromfile(os.path.join(tmp, "vals.bin"), dtype=np.complex128))
        for i in range(n_mut):
            out["mutated"].append(np.fromfile(os.path.join(tmp, "mut%d.bin" % i), dtype=np.complex128))
        for line in res.stdout.splitlines():
            if line.startswith("TIMING ops"):
                out["seconds"] = float(line.split()[2])
        return out
