// examples/main.cpp -- the reference's timing demo (main.cpp:19-40: a 26-qubit state vector, one 4-target random gate
// on {0,6,4,2}, comm_synch() + std::chrono around the call) written against the drop-in headers. The two helpers the
// reference takes from its test utilities (getRandomMatrix, rootNodePrint) are inlined here.
//
// build: g++ -std=c++17 -O2 -Iinclude -I<pkg>/host examples/main.cpp -o examples/main -L<pkg> -ldfsa_b200 -Wl,-rpath,<pkg>
// run:   ./examples/main [numQubits]          (1 rank)      DFSA_NP=2 ./examples/main      (2 ranks, one per GPU)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "distributed_densitymatrix.hpp"
#include "distributed_statevector.hpp"

static AmpMatrix randomMatrix(Index dim) {
    AmpMatrix m = getZeroMatrix(dim);
    for (Index r = 0; r < dim; r++)
        for (Index c = 0; c < dim; c++) {
            Real a = (std::rand() + 1.0) / (RAND_MAX + 1.0), b = std::rand() / Real(RAND_MAX);
            m[r][c] = Amp(std::sqrt(-2 * std::log(a)) * std::cos(2 * 3.14159265 * b), std::sqrt(-2 * std::log(a)) * std::sin(2 * 3.14159265 * b));
        }
    return m;
}

static void rootPrint(const std::string& msg) {
    comm_synch();
    if (comm_getRank() == 0) std::printf("%s\n", msg.c_str());
    std::fflush(stdout);
    comm_synch();
}

int main(int argc, char** argv) {
    comm_init();

    Nat numQubits = argc > 1 ? Nat(std::atoi(argv[1])) : 26;
    StateVector state = StateVector(numQubits);
    state.setHashAmps(1);

    NatArray targets = {0, 6, 4, 2};
    AmpMatrix matrix = randomMatrix(powerOf2(Nat(targets.size())));

    distributed_statevector_manyTargGate(state, targets, matrix);     // warm-up (first launch pays module load)
    comm_synch();
    auto start = std::chrono::high_resolution_clock::now();

    distributed_statevector_manyTargGate(state, targets, matrix);

    comm_synch();
    auto stop = std::chrono::high_resolution_clock::now();
    auto dur = std::chrono::duration_cast<std::chrono::microseconds>(stop - start).count();
    rootPrint("done in " + std::to_string(dur) + " microseconds on " + std::to_string(comm_getNumNodes()) + " rank(s), transport " + dfsa_comm_transport());

    comm_end();
    return 0;
}
