// fuzz.cpp -- TEST INFRASTRUCTURE: random circuits through the REAL drop-in host layer (host/*.hpp: local-vs-exchange decisions,
// relocation plans, the lazy qubit layout, swap-in of rank-bit qubits, the deferred gate queue, layout restores, partialTrace's
// reordering) at DFSA_NP = 1...16 ranks, on the CPU stand-in of the C-ABI (dfsa_hostsim.cpp), against a dense ground truth that
// applies every operation to the whole 2^n vector by its textbook definition and knows nothing about ranks or layouts.
//
//   DFSA_NP=8 tests/hostsim/_build/fuzz [trials=60] [seed=1]          exit code 0 = every comparison on every rank agreed
//
// Tolerance 1e-10 * max(1, max|truth|): the stand-in and the truth differ by rounding only. The gates are random unitaries so that
// long circuits keep O(1) amplitudes.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "distributed_densitymatrix.hpp"

extern "C" unsigned long long hostsim_call_count(const char* entry);      // dfsa_hostsim.cpp

namespace {

struct Rng {
    unsigned long long s;
    unsigned long long next() { s += 0x9E3779B97F4A7C15ULL; unsigned long long z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
    double uni() { return double(next() >> 11) / double(1ULL << 53); }
    Nat below(Nat n) { return Nat(next() % n); }
    Nat between(Nat lo, Nat hi) { return lo + below(hi - lo + 1); }
    Amp gauss() { const double a = uni() + 1e-300, b = uni(); const double r = std::sqrt(-2 * std::log(a)); return Amp(r * std::cos(6.283185307179586 * b), r * std::sin(6.283185307179586 * b)); }
    NatArray distinct(Nat count, Nat n) {
        NatArray all(n);
        for (Nat q = 0; q < n; q++) all[q] = q;
        for (Nat i = 0; i < count; i++) std::swap(all[i], all[i + below(n - i)]);
        all.resize(count);
        return all;
    }
};

AmpMatrix randomMatrix(Rng& rng, Index dim) {
    AmpMatrix m = getZeroMatrix(dim);
    for (auto& row : m) for (Amp& e : row) e = rng.gauss();
    return m;
}
AmpMatrix randomUnitary(Rng& rng, Index dim) {              // Gram-Schmidt on the rows of a Gaussian matrix
    AmpMatrix m = randomMatrix(rng, dim);
    for (Index r = 0; r < dim; r++) {
        for (Index p = 0; p < r; p++) {
            Amp dot(0, 0);
            for (Index c = 0; c < dim; c++) dot += std::conj(m[p][c]) * m[r][c];
            for (Index c = 0; c < dim; c++) m[r][c] -= dot * m[p][c];
        }
        double norm = 0;
        for (Index c = 0; c < dim; c++) norm += std::norm(m[r][c]);
        norm = std::sqrt(norm);
        for (Index c = 0; c < dim; c++) m[r][c] /= norm;
    }
    return m;
}

// ---- dense ground truth on the whole vector (bit q of the index = qubit q) ------------------------------------------------
void truthManyTarg(AmpArray& v, const NatArray& targets, const AmpMatrix& G) {
    const Index D = Index(1) << targets.size();
    Index mask = 0;
    for (Nat t : targets) mask |= Index(1) << t;
    AmpArray in(D);
    for (Index base = 0; base < v.size(); base++) {
        if (base & mask) continue;
        for (Index c = 0; c < D; c++) in[c] = v[setBits(base, targets, c)];
        for (Index r = 0; r < D; r++) {
            Amp acc(0, 0);
            for (Index c = 0; c < D; c++) acc += G[r][c] * in[c];
            v[setBits(base, targets, r)] = acc;
        }
    }
}
void truthCtrlOneTarg(AmpArray& v, const NatArray& ctrls, Nat target, const AmpMatrix& g) {
    const Index t = Index(1) << target;
    for (Index i = 0; i < v.size(); i++) {
        if ((i & t) || !allBitsAreOne(i, ctrls)) continue;
        const Amp a0 = v[i], a1 = v[i | t];
        v[i] = g[0][0] * a0 + g[0][1] * a1;
        v[i | t] = g[1][0] * a0 + g[1][1] * a1;
    }
}
void truthSwap(AmpArray& v, Nat q1, Nat q2) {
    for (Index i = 0; i < v.size(); i++)
        if (getBit(i, q1) && !getBit(i, q2)) std::swap(v[i], v[flipTwoBits(i, q1, q2)]);
}
// P v for P = tensor of 2x2 Pauli matrices (element by element from the matrices; `conjugated`: P* instead of P)
AmpArray truthPauliApplied(const AmpArray& v, const NatArray& targets, const NatArray& paulis, bool conjugated) {
    static const Amp mats[4][2][2] = {{{1, 0}, {0, 1}}, {{0, 1}, {1, 0}}, {{0, Amp(0, -1)}, {Amp(0, 1), 0}}, {{1, 0}, {0, -1}}};
    AmpArray out(v.size());
    for (Index i = 0; i < v.size(); i++) {
        Index j = i;
        Amp elem(1, 0);
        for (std::size_t k = 0; k < targets.size(); k++) {
            const Nat bi = getBit(i, targets[k]);
            const Nat bj = (paulis[k] == X || paulis[k] == Y) ? !bi : bi;
            if (bj != bi) j = flipBit(j, targets[k]);
            elem *= conjugated ? std::conj(mats[paulis[k]][bi][bj]) : mats[paulis[k]][bi][bj];
        }
        out[i] = elem * v[j];
    }
    return out;
}
void truthPauliGadget(AmpArray& v, const NatArray& targets, const NatArray& paulis, Real theta, bool conjugated) {
    const AmpArray pv = truthPauliApplied(v, targets, paulis, conjugated);     // exp(i theta P) = cos + i sin P; its conjugate = cos - i sin P*
    const Amp fac = conjugated ? Amp(0, -std::sin(theta)) : Amp(0, std::sin(theta));
    for (Index i = 0; i < v.size(); i++) v[i] = std::cos(theta) * v[i] + fac * pv[i];
}
void truthPhaseGadget(AmpArray& v, const NatArray& targets, Real theta) {      // exp(i theta Z...Z)
    const Index mask = getBitMask(targets);
    for (Index i = 0; i < v.size(); i++) v[i] *= getBitMaskParity(i & mask) ? Amp(std::cos(theta), -std::sin(theta)) : Amp(std::cos(theta), std::sin(theta));
}
NatArray plusN(const NatArray& qubits, Nat N) { NatArray out = qubits; for (Nat& q : out) q += N; return out; }

struct Stats { unsigned long long checks = 0, failures = 0; };

bool closeEnough(const AmpArray& got, const AmpArray& truth, const std::string& what) {
    if (got.size() != truth.size()) { std::fprintf(stderr, "[rank %u] %s: size %zu != %zu\n", comm_getRank(), what.c_str(), got.size(), truth.size()); return false; }
    double scale = 1, worst = 0;
    Index at = 0;
    for (const Amp& a : truth) scale = std::max(scale, std::max(std::abs(a.real()), std::abs(a.imag())));
    for (Index i = 0; i < truth.size(); i++) {
        const double d = std::max(std::abs(got[i].real() - truth[i].real()), std::abs(got[i].imag() - truth[i].imag()));
        if (!(d <= worst)) { worst = d; at = i; }
    }
    if (worst <= 1e-10 * scale) return true;
    std::fprintf(stderr, "[rank %u] MISMATCH %s: |delta| = %g at index %llu (got %g%+gi, truth %g%+gi)\n", comm_getRank(), what.c_str(), worst, at,
                 got[at].real(), got[at].imag(), truth[at].real(), truth[at].imag());
    return false;
}

std::string show(const NatArray& a) { std::string s = "["; for (Nat x : a) s += std::to_string(x) + ","; return s + "]"; }

// ---- one state-vector trial -----------------------------------------------------------------------------------------------
void svOps(Rng& rng, StateVector& psi, AmpArray& truth, Nat n, Nat numOps, Stats& st, std::string& log, const std::string& tag) {
    const Nat L = Nat(psi.logNumAmpsPerNode);
    for (Nat op = 0; op < numOps; op++) {
        const Nat kind = rng.below(11);
        if (kind == 10) {
            // a burst of one-target gates (what the gate queue and its swap-in planning are for): either a sweep over consecutive
            // qubits starting anywhere, or random targets; each gate with 0-3 controls
            const Nat count = rng.below(25) == 0 ? rng.between(250, 600) : rng.between(3, 40), start = rng.below(n);      // (now and then past the queue's 256-gate limit)
            const bool sweep = rng.below(2) == 0;
            log += sweep ? " sweep(" + std::to_string(start) + "x" + std::to_string(count) + ")" : " burst(" + std::to_string(count) + ")";
            for (Nat b = 0; b < count; b++) {
                const Nat t = sweep ? (start + b / 2) % n : rng.below(n);
                NatArray ctrls;
                for (Nat q : rng.distinct(rng.below(std::min<Nat>(4, n)), n)) if (q != t) ctrls.push_back(q);
                const AmpMatrix g = randomUnitary(rng, 2);
                if (ctrls.empty()) distributed_statevector_oneTargGate(psi, t, g);
                else distributed_statevector_manyCtrlOneTargGate(psi, ctrls, t, g);
                truthCtrlOneTarg(truth, ctrls, t, g);
            }
        } else if (kind == 0) {
            const Nat t = rng.below(n);
            const AmpMatrix g = randomUnitary(rng, 2);
            log += " one(" + std::to_string(t) + ")";
            distributed_statevector_oneTargGate(psi, t, g);
            truthCtrlOneTarg(truth, {}, t, g);
        } else if (kind <= 2) {
            const NatArray q = rng.distinct(rng.between(2, std::min<Nat>(4, n)), n);
            const NatArray ctrls(q.begin() + 1, q.end());
            const AmpMatrix g = randomUnitary(rng, 2);
            log += " ctrl(" + show(ctrls) + "," + std::to_string(q[0]) + ")";
            distributed_statevector_manyCtrlOneTargGate(psi, ctrls, q[0], g);
            truthCtrlOneTarg(truth, ctrls, q[0], g);
        } else if (kind == 3) {
            const NatArray q = rng.distinct(2, n);
            log += " swap(" + std::to_string(q[0]) + "," + std::to_string(q[1]) + ")";
            distributed_statevector_swapGate(psi, q[0], q[1]);
            truthSwap(truth, q[0], q[1]);
        } else if (kind <= 5) {
            const NatArray targets = rng.distinct(rng.between(1, std::min<Nat>(4, L)), n);
            const AmpMatrix G = randomUnitary(rng, Index(1) << targets.size());
            log += " many(" + show(targets) + ")";
            distributed_statevector_manyTargGate(psi, targets, G);
            truthManyTarg(truth, targets, G);
        } else if (kind == 6) {
            const NatArray targets = rng.distinct(rng.between(1, std::min<Nat>(6, n)), n);
            NatArray paulis(targets.size());
            for (Nat& p : paulis) p = rng.below(4);
            log += " ptensor(" + show(targets) + show(paulis) + ")";
            distributed_statevector_pauliTensor(psi, targets, paulis);
            truth = truthPauliApplied(truth, targets, paulis, false);
        } else if (kind == 7) {
            const NatArray targets = rng.distinct(rng.between(1, std::min<Nat>(6, n)), n);
            NatArray paulis(targets.size());
            for (Nat& p : paulis) p = rng.below(4);
            const Real theta = (rng.uni() - 0.5) * 6.0;
            log += " pgadget(" + show(targets) + show(paulis) + ")";
            distributed_statevector_pauliGadget(psi, targets, paulis, theta);
            truthPauliGadget(truth, targets, paulis, theta, false);
        } else if (kind == 8) {
            const NatArray targets = rng.distinct(rng.between(1, n), n);
            const Real theta = (rng.uni() - 0.5) * 6.0;
            log += " phase(" + show(targets) + ")";
            distributed_statevector_phaseGadget(psi, targets, theta);
            truthPhaseGadget(truth, targets, theta);
        } else {
            const Nat which = rng.below(3);
            if (which == 0) { log += " synch"; comm_synch(); }
            else {
                log += " check";
                st.checks++;
                if (!closeEnough(psi.getAllVecAmps(), truth, tag + " after:" + log)) st.failures++;
            }
        }
    }
}

void svTrial(Rng& rng, Nat k, Stats& st, const std::string& tag) {
    const Nat n = rng.between(std::max<Nat>(k + 2, 4), std::max<Nat>(k + 2, 4) + 5);
    StateVector psi(n);
    AmpArray truth(Index(1) << n);
    for (Amp& a : truth) a = rng.gauss();
    psi.setAllVecAmps(truth);
    std::string log = " n=" + std::to_string(n);
    svOps(rng, psi, truth, n, rng.between(10, 50), st, log, tag);
    st.checks++;
    if (!closeEnough(psi.getAllVecAmps(), truth, tag + " final:" + log)) st.failures++;
}

// ---- one density-matrix trial: Choi vector of 2N bits, flat = 2^N * col + row ------------------------------------------------
void dmTrial(Rng& rng, Nat k, Stats& st, const std::string& tag) {
    const Nat N = rng.between(std::max<Nat>(k + 1, 2), std::max<Nat>(k + 1, 2) + 2);
    DensityMatrix rho(N);
    const Nat L = Nat(rho.logNumAmpsPerNode);
    AmpArray truth(Index(1) << (2 * N));
    for (Amp& a : truth) a = rng.gauss();
    rho.setAllVecAmps(truth);
    std::string log = " N=" + std::to_string(N);
    const Nat numOps = rng.between(8, 30);
    for (Nat op = 0; op < numOps; op++) {
        const Nat kind = rng.below(12);
        if (kind <= 1) {
            const NatArray targets = rng.distinct(rng.between(1, std::min<Nat>(3, std::min<Nat>(N, L / 2))), N);
            const AmpMatrix G = randomUnitary(rng, Index(1) << targets.size());
            log += " dm_many(" + show(targets) + ")";
            distributed_densitymatrix_manyTargGate(rho, targets, G);
            truthManyTarg(truth, targets, G);
            truthManyTarg(truth, plusN(targets, N), getConjugateMatrix(G));
        } else if (kind == 2 && N >= 2) {
            const NatArray q = rng.distinct(2, N);
            log += " dm_swap(" + std::to_string(q[0]) + "," + std::to_string(q[1]) + ")";
            distributed_densitymatrix_swapGate(rho, q[0], q[1]);
            truthSwap(truth, q[0], q[1]);
            truthSwap(truth, q[0] + N, q[1] + N);
        } else if (kind == 3) {
            const NatArray targets = rng.distinct(rng.between(1, N), N);
            NatArray paulis(targets.size());
            for (Nat& p : paulis) p = rng.below(4);
            log += " dm_ptensor(" + show(targets) + show(paulis) + ")";
            distributed_densitymatrix_pauliTensor(rho, targets, paulis);
            truth = truthPauliApplied(truthPauliApplied(truth, targets, paulis, false), plusN(targets, N), paulis, true);
        } else if (kind == 4) {
            const NatArray targets = rng.distinct(rng.between(1, N), N);
            NatArray paulis(targets.size());
            for (Nat& p : paulis) p = rng.below(4);
            const Real theta = (rng.uni() - 0.5) * 6.0;
            log += " dm_pgadget(" + show(targets) + show(paulis) + ")";
            distributed_densitymatrix_pauliGadget(rho, targets, paulis, theta);
            truthPauliGadget(truth, targets, paulis, theta, false);
            truthPauliGadget(truth, plusN(targets, N), paulis, theta, true);
        } else if (kind == 5) {
            const NatArray targets = rng.distinct(rng.between(1, N), N);
            const Real theta = (rng.uni() - 0.5) * 6.0;
            log += " dm_phase(" + show(targets) + ")";
            distributed_densitymatrix_phaseGadget(rho, targets, theta);
            truthPhaseGadget(truth, targets, theta);
            truthPhaseGadget(truth, plusN(targets, N), -theta);
        } else if (kind == 6) {
            const NatArray targets = rng.distinct(rng.between(1, std::min<Nat>(2, std::min<Nat>(N, L / 2))), N);
            const Index d = Index(1) << targets.size();
            MatrixArray ops(rng.between(1, 3));
            for (AmpMatrix& K : ops) K = Amp(1.0 / std::sqrt(double(ops.size())), 0) * randomUnitary(rng, d);
            log += " kraus(" + show(targets) + "x" + std::to_string(ops.size()) + ")";
            distributed_densitymatrix_krausMap(rho, ops, targets);
            AmpArray sum(truth.size(), Amp(0, 0));                  // rho' = sum_K K rho K^dagger
            for (const AmpMatrix& K : ops) {
                AmpArray term = truth;
                truthManyTarg(term, targets, K);
                truthManyTarg(term, plusN(targets, N), getConjugateMatrix(K));
                for (Index i = 0; i < sum.size(); i++) sum[i] += term[i];
            }
            truth = sum;
        } else if (kind == 7) {
            const Nat q = rng.below(N);
            const Real p = rng.uni() * 0.5;
            log += " deph1(" + std::to_string(q) + ")";
            distributed_densitymatrix_oneQubitDephasing(rho, q, p);
            for (Index i = 0; i < truth.size(); i++) if (getBit(i, q) != getBit(i, q + N)) truth[i] *= 1 - 2 * p;
        } else if (kind == 8 && N >= 2) {
            const NatArray q = rng.distinct(2, N);
            const Real p = rng.uni() * 0.5;
            log += " deph2(" + std::to_string(q[0]) + "," + std::to_string(q[1]) + ")";
            distributed_densitymatrix_twoQubitDephasing(rho, q[0], q[1], p);
            for (Index i = 0; i < truth.size(); i++)
                if (getBit(i, q[0]) != getBit(i, q[0] + N) || getBit(i, q[1]) != getBit(i, q[1] + N)) truth[i] *= 1 - 4 * p / 3;
        } else if (kind == 9) {
            // the state-vector API on the Choi vector's 2N bits (a DensityMatrix is a StateVector): one-target gates on rank bits swap
            // the qubit into the shard, and whatever comes next has to cope with that layout
            svOps(rng, rho, truth, 2 * N, rng.between(1, 4), st, log, tag);
        } else if (kind == 10) {
            log += " check";
            st.checks++;
            if (!closeEnough(rho.getAllVecAmps(), truth, tag + " after:" + log)) st.failures++;
        }
    }
    st.checks++;
    if (!closeEnough(rho.getAllVecAmps(), truth, tag + " final:" + log)) st.failures++;
    // partialTrace of 1..N-max(k,1) qubits: out[r', c'] = sum_k in[r' with k inserted, c' with k inserted]
    const Nat maxTraced = N - std::max<Nat>(k, 1);
    if (maxTraced >= 1 && rng.below(3) != 0) {
        NatArray targets = rng.distinct(rng.between(1, maxTraced), N);
        log += " ptrace(" + show(targets) + ")";
        DensityMatrix out = distributed_densitymatrix_partialTrace(rho, targets);
        std::sort(targets.begin(), targets.end());
        const Nat t = Nat(targets.size()), M = N - t;
        AmpArray expect(Index(1) << (2 * M), Amp(0, 0));
        for (Index c = 0; c < (Index(1) << M); c++)
            for (Index r = 0; r < (Index(1) << M); r++)
                for (Index kk = 0; kk < (Index(1) << t); kk++) {
                    const Index rr = setBits(insertBits(r, targets, 0), targets, kk), cc = setBits(insertBits(c, targets, 0), targets, kk);
                    expect[(c << M) | r] += truth[(cc << N) | rr];
                }
        st.checks++;
        if (!closeEnough(out.getAllVecAmps(), expect, tag + " partialTrace:" + log)) st.failures++;
    }
}

}  // namespace

int main(int argc, char** argv) {
    const unsigned trials = argc > 1 ? unsigned(std::atoi(argv[1])) : 60;
    const unsigned long long seed = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 1;
    comm_init();
    const Nat P = comm_getNumNodes(), k = logBase2(P);
    Stats st;
    for (unsigned trial = 0; trial < trials; trial++) {
        Rng rng{seed * 1000003ULL + trial};
        const std::string tag = "P=" + std::to_string(P) + " seed=" + std::to_string(seed) + " trial=" + std::to_string(trial);
        if (trial % 2 == 0) svTrial(rng, k, st, tag + " sv");
        else dmTrial(rng, k, st, tag + " dm");
    }
    Amp total(double(st.failures), double(st.checks));
    comm_reduceAmp(total);
    if (comm_getRank() == 0)
        std::printf("hostsim fuzz: P=%u lazy_layout=%d gate_fusion=%d trials=%u comparisons=%.0f failures=%.0f\n", P, int(dfsa_detail::lazyLayoutEnabled()),
                    int(StateVector::gateFusionEnabled()), trials, total.imag(), total.real());
    // how often the ranks (summed) entered each C-ABI function: lets the test assert that the paths it means to exercise were taken
    std::string calls = "calls, all ranks:";
    for (const char* e : {"dfsa_k_gateSequence", "dfsa_k_ctrlOneTarg", "dfsa_k_manyTarg", "dfsa_k_krausMap", "dfsa_k_pauli", "dfsa_k_swap", "dfsa_k_partialTrace", "dfsa_xk_relocate",
                          "dfsa_xk_swapSuffixPrefix", "dfsa_xk_exchangeCombine", "dfsa_xk_ctrlPrefixTarg", "dfsa_xk_exchangePauliCombine", "dfsa_x_exchange"}) {
        Amp count(double(hostsim_call_count(e)), 0);
        comm_reduceAmp(count);
        calls += " " + std::string(e + 5) + "=" + std::to_string((unsigned long long)count.real());
    }
    if (comm_getRank() == 0) std::printf("%s\n", calls.c_str());
    comm_end();
    return total.real() == 0 ? 0 : 1;
}
