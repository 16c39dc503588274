/*
 * ref_driver.cpp -- runs an "op script" through the UNMODIFIED reference API
 * (src/distributed_statevector.hpp, src/distributed_densitymatrix.hpp) on the host CPU.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (never linked into the product). It is compiled by
 * oracle/build_ref.sh against a build-time copy of /root/reference/src in which the one-token
 * setBit bug (src/bit_maths.hpp:66) is patched, and against oracle/mpi_shim/mpi.h.
 * Output binary: oracle/_ref/ref_driver. Uses: (1) golden vectors for tests/golden/,
 * (2) live cross-check of the C restatement in oracle/dfsa_oracle.c, (3) the CPU baseline /
 * `bench.py --impl reference` timing with the reference's own OpenMP loops.
 *
 * Script grammar (whitespace separated; reals may be decimal or C99 hex floats):
 *   state sv|dm N | init file PATH | init hash SEED | dump PATH | dumpvals PATH
 *   tic | toc LABEL | <op> ...          (op names = reference API names, see run() below)
 * Launch: SHIM_NP=P OMP_NUM_THREADS=T oracle/_ref/ref_driver script.txt
 */
#include "types.hpp"
#include "states.hpp"
#include "distributed_statevector.hpp"
#include "distributed_densitymatrix.hpp"

#include <chrono>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <fcntl.h>
#include <unistd.h>

static uint64_t splitmix64(uint64_t seed, uint64_t k) {
    uint64_t z = seed + (k + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static double hashReal(uint64_t seed, uint64_t k) {
    return (double)(splitmix64(seed, k) >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

struct Reader {
    std::ifstream in;
    explicit Reader(const char* path) : in(path) { if (!in) { fprintf(stderr, "cannot open %s\n", path); exit(2); } }
    bool word(std::string& w) { return (bool)(in >> w); }
    std::string str() { std::string w; if (!(in >> w)) { fprintf(stderr, "script truncated\n"); exit(2); } return w; }
    Nat nat() { return (Nat) std::stoul(str()); }
    Real real() { return strtod(str().c_str(), nullptr); }
    Amp amp() { Real re = real(); Real im = real(); return Amp(re, im); }
    NatArray nats(Nat n) { NatArray a(n); for (Nat& x : a) x = nat(); return a; }
    AmpMatrix matrix(Index dim) {
        AmpMatrix m = getZeroMatrix(dim);
        for (Index r = 0; r < dim; r++) for (Index c = 0; c < dim; c++) m[r][c] = amp();
        return m;
    }
};

static void pwriteAll(int fd, const void* buf, size_t bytes, off_t off) {
    const char* p = (const char*)buf;
    while (bytes > 0) {
        ssize_t n = pwrite(fd, p, bytes, off);
        if (n <= 0) { perror("pwrite"); exit(2); }
        p += n; bytes -= n; off += n;
    }
}

static void dumpState(StateVector& psi, const std::string& path) {
    comm_synch();
    if (psi.rank == 0) { int fd = open(path.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644); close(fd); }
    comm_synch();
    int fd = open(path.c_str(), O_WRONLY);
    if (fd < 0) { perror("open dump"); exit(2); }
    pwriteAll(fd, psi.amps.data(), psi.numAmpsPerNode * sizeof(Amp), (off_t)(psi.rank * psi.numAmpsPerNode * sizeof(Amp)));
    close(fd);
    comm_synch();
}

static void run(const char* scriptPath) {
    Reader rd(scriptPath);
    std::unique_ptr<DensityMatrix> rho;   /* DensityMatrix derives from StateVector (src/states.hpp:53) */
    std::unique_ptr<StateVector> psiOnly;
    StateVector* psi = nullptr;
    std::vector<Amp> values;
    std::chrono::high_resolution_clock::time_point t0;

    auto needDM = [&]() -> DensityMatrix& { if (!rho) { fprintf(stderr, "op needs a dm state\n"); exit(2); } return *rho; };
    auto needSV = [&]() -> StateVector& { if (!psi) { fprintf(stderr, "op needs a state\n"); exit(2); } return *psi; };

    std::string w;
    while (rd.word(w)) {
        if (w == "state") {
            std::string kind = rd.str(); Nat n = rd.nat();
            rho.reset(); psiOnly.reset();
            if (kind == "dm") { rho.reset(new DensityMatrix(n)); psi = rho.get(); }
            else { psiOnly.reset(new StateVector(n)); psi = psiOnly.get(); }
        } else if (w == "init") {
            std::string how = rd.str();
            StateVector& s = needSV();
            Index first = (Index)s.rank * s.numAmpsPerNode;
            if (how == "file") {
                std::string path = rd.str();
                FILE* f = fopen(path.c_str(), "rb");
                if (!f) { perror("init file"); exit(2); }
                fseeko(f, (off_t)(first * sizeof(Amp)), SEEK_SET);
                if (fread(s.amps.data(), sizeof(Amp), s.numAmpsPerNode, f) != s.numAmpsPerNode) { fprintf(stderr, "short init file\n"); exit(2); }
                fclose(f);
            } else {
                uint64_t seed = std::stoull(rd.str());
                #pragma omp parallel for
                for (Index j = 0; j < s.numAmpsPerNode; j++)
                    s.amps[j] = Amp(hashReal(seed, 2 * (first + j)), hashReal(seed, 2 * (first + j) + 1));
            }
        } else if (w == "dump") {
            dumpState(needSV(), rd.str());
        } else if (w == "dumpvals") {
            std::string path = rd.str();
            if (comm_getRank() == 0) {
                FILE* f = fopen(path.c_str(), "wb");
                fwrite(values.data(), sizeof(Amp), values.size(), f);
                fclose(f);
            }
        } else if (w == "tic") {
            comm_synch();
            t0 = std::chrono::high_resolution_clock::now();
        } else if (w == "toc") {
            std::string label = rd.str();
            comm_synch();
            double s = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
            if (comm_getRank() == 0) { printf("TIMING %s %.9f\n", label.c_str(), s); fflush(stdout); }
        }
        /* ---- state-vector API (src/distributed_statevector.hpp:18,81,109,190,279,287,295) ---- */
        else if (w == "sv_oneTargGate") {
            Nat t = rd.nat(); AmpMatrix g = rd.matrix(2);
            distributed_statevector_oneTargGate(needSV(), t, g);
        } else if (w == "sv_manyCtrlOneTargGate") {
            Nat nc = rd.nat(); NatArray c = rd.nats(nc); Nat t = rd.nat(); AmpMatrix g = rd.matrix(2);
            distributed_statevector_manyCtrlOneTargGate(needSV(), c, t, g);
        } else if (w == "sv_swapGate") {
            Nat a = rd.nat(); Nat b = rd.nat();
            distributed_statevector_swapGate(needSV(), a, b);
        } else if (w == "sv_manyTargGate") {
            Nat nt = rd.nat(); NatArray t = rd.nats(nt); AmpMatrix g = rd.matrix(powerOf2(nt));
            distributed_statevector_manyTargGate(needSV(), t, g);
        } else if (w == "sv_pauliTensor") {
            Nat nt = rd.nat(); NatArray t = rd.nats(nt); NatArray p = rd.nats(nt);
            distributed_statevector_pauliTensor(needSV(), t, p);
        } else if (w == "sv_pauliGadget") {
            Nat nt = rd.nat(); NatArray t = rd.nats(nt); NatArray p = rd.nats(nt); Real th = rd.real();
            distributed_statevector_pauliGadget(needSV(), t, p, th);
        } else if (w == "sv_phaseGadget") {
            Nat nt = rd.nat(); NatArray t = rd.nats(nt); Real th = rd.real();
            distributed_statevector_phaseGadget(needSV(), t, th);
        }
        /* ---- density-matrix API (src/distributed_densitymatrix.hpp:15..347) ---- */
        else if (w == "dm_manyTargGate") {
            Nat nt = rd.nat(); NatArray t = rd.nats(nt); AmpMatrix g = rd.matrix(powerOf2(nt));
            distributed_densitymatrix_manyTargGate(needDM(), t, g);
        } else if (w == "dm_swapGate") {
            Nat a = rd.nat(); Nat b = rd.nat();
            distributed_densitymatrix_swapGate(needDM(), a, b);
        } else if (w == "dm_pauliTensor") {
            Nat nt = rd.nat(); NatArray t = rd.nats(nt); NatArray p = rd.nats(nt);
            distributed_densitymatrix_pauliTensor(needDM(), t, p);
        } else if (w == "dm_pauliGadget") {
            Nat nt = rd.nat(); NatArray t = rd.nats(nt); NatArray p = rd.nats(nt); Real th = rd.real();
            distributed_densitymatrix_pauliGadget(needDM(), t, p, th);
        } else if (w == "dm_phaseGadget") {
            Nat nt = rd.nat(); NatArray t = rd.nats(nt); Real th = rd.real();
            distributed_densitymatrix_phaseGadget(needDM(), t, th);
        } else if (w == "dm_krausMap") {
            Nat nt = rd.nat(); NatArray t = rd.nats(nt); Nat nk = rd.nat();
            MatrixArray ops(nk);
            for (AmpMatrix& m : ops) m = rd.matrix(powerOf2(nt));
            distributed_densitymatrix_krausMap(needDM(), ops, t);
        } else if (w == "dm_oneQubitDephasing") {
            Nat q = rd.nat(); Real p = rd.real();
            distributed_densitymatrix_oneQubitDephasing(needDM(), q, p);
        } else if (w == "dm_twoQubitDephasing") {
            Nat a = rd.nat(); Nat b = rd.nat(); Real p = rd.real();
            distributed_densitymatrix_twoQubitDephasing(needDM(), a, b, p);
        } else if (w == "dm_oneQubitDepolarising") {
            Nat q = rd.nat(); Real p = rd.real();
            distributed_densitymatrix_oneQubitDepolarising(needDM(), q, p);
        } else if (w == "dm_twoQubitDepolarising") {
            Nat a = rd.nat(); Nat b = rd.nat(); Real p = rd.real();
            distributed_densitymatrix_twoQubitDepolarising(needDM(), a, b, p);
        } else if (w == "dm_damping") {
            Nat q = rd.nat(); Real p = rd.real();
            distributed_densitymatrix_damping(needDM(), q, p);
        } else if (w == "dm_expecPauliString") {
            Nat nterms = rd.nat();
            RealArray coeffs(nterms); for (Real& c : coeffs) c = rd.real();
            NatArray paulis = rd.nats(nterms * needDM().numQubits);
            values.push_back(distributed_densitymatrix_expecPauliString(needDM(), coeffs, paulis));
        } else if (w == "dm_partialTrace") {
            Nat nt = rd.nat(); NatArray t = rd.nats(nt); std::string mutatedPath = rd.str();
            DensityMatrix out = distributed_densitymatrix_partialTrace(needDM(), t);
            if (mutatedPath != "-") dumpState(needDM(), mutatedPath);   /* the reference mutates its input */
            rho.reset(new DensityMatrix(std::move(out))); psi = rho.get();
        } else {
            fprintf(stderr, "ref_driver: unknown word '%s'\n", w.c_str()); exit(2);
        }
    }
}

int main(int argc, char** argv) {
    if (argc != 2) { fprintf(stderr, "usage: SHIM_NP=P %s script.txt\n", argv[0]); return 2; }
    comm_init();
    run(argv[1]);
    comm_end();
    return 0;
}
