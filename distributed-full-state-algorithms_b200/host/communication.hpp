// communication.hpp -- the communication environment of the drop-in API, on NCCL / NVLink instead of MPI.
// Same entry points as the reference's src/communication.hpp (:16 comm_init, :24 comm_end, :30 comm_getRank,
// :37 comm_getNumNodes, :44 comm_synch, :100/:110 comm_exchangeArrays, :141 comm_asynchSendArray,
// :161 comm_receiveArray, :172 comm_reduceAmp). The array arguments are the device-resident DeviceAmpArray
// handles of states.hpp instead of std::vector; messages are not split into 2^30-amplitude chunks because
// NCCL has no such count limit. One process per GPU; launch with torchrun-style RANK/WORLD_SIZE, or set
// DFSA_NP=P to have comm_init() fork the ranks itself (replaces `mpirun -np P`).
#pragma once

#include <cstdio>
#include <cstdlib>

#include "dfsa_b200.h"
#include "types.hpp"

// every C-ABI failure is fatal, like the reference's asserts (states.hpp:35, distributed_statevector.hpp:191)
#define DFSA_CHECK(call)                                                                                   \
    do {                                                                                                   \
        int dfsa_rc_ = (call);                                                                             \
        if (dfsa_rc_ != 0) {                                                                               \
            std::fprintf(stderr, "%s:%d: %s failed (%d): %s\n", __FILE__, __LINE__, #call, dfsa_rc_, dfsa_last_error()); \
            std::abort();                                                                                  \
        }                                                                                                  \
    } while (0)

// a view of one of a state's two device arrays (psi.amps / psi.buffer)
struct DeviceAmpArray {
    dfsa_state* owner = nullptr;
    int which = DFSA_AMPS;
    Index size() const { return owner ? dfsa_state_num_amps_per_node(owner) : 0; }
    Amp* data() const { return reinterpret_cast<Amp*>(dfsa_state_ptr(owner, which)); }   // DEVICE pointer
};

// Deferred work: a state may hold one-target gates it has not launched yet (states.hpp: gateQueue, fused into shared passes over
// HBM). comm_synch() -- the reference's "everything before this point has happened" (communication.hpp:44), which is what
// main.cpp-style timing brackets with -- launches them first; states.hpp installs the hook.
namespace dfsa_detail {
inline void (*&flushAllStatesHook())() { static void (*hook)() = nullptr; return hook; }
}

static inline void comm_init() { DFSA_CHECK(dfsa_comm_init()); }
// Joining a job whose ranks do not share a parent process or a node-local rendezvous (srun / mpirun wrappers that set neither
// MASTER_PORT nor DFSA_JOB_ID): rank 0 obtains a 128-byte id with dfsa_comm_get_unique_id, the launcher's own mechanism
// broadcasts it, every rank calls this instead of comm_init(). Exchanges then run on the staged NCCL transport.
static inline void comm_initWithId(Nat rank, Nat numRanks, const void* uniqueId128, int device = -1) {
    DFSA_CHECK(dfsa_comm_init_with_id(int(rank), int(numRanks), uniqueId128, device));
}
static inline void comm_end() { DFSA_CHECK(dfsa_comm_finalize()); }
static inline Nat comm_getRank() { return Nat(dfsa_comm_rank()); }
static inline Nat comm_getNumNodes() { return Nat(dfsa_comm_size()); }
static inline void comm_synch() {
    if (dfsa_detail::flushAllStatesHook()) dfsa_detail::flushAllStatesHook()();
    DFSA_CHECK(dfsa_comm_barrier());
}

static inline void comm_exchangeArrays(DeviceAmpArray& toSend, Index toSendStartInd, DeviceAmpArray& toReceive, Index toReceiveStartInd,
                                       Index numAmpsToExchange, Nat pairRank) {
    DFSA_CHECK(dfsa_x_exchange(toSend.owner, toSend.which, toSendStartInd, toReceive.which, toReceiveStartInd, numAmpsToExchange, int(pairRank)));
}

static inline void comm_exchangeArrays(DeviceAmpArray& toSend, DeviceAmpArray& toReceive, Nat pairRank) {
    comm_exchangeArrays(toSend, 0, toReceive, 0, toSend.size(), pairRank);
}

// one-directional pair (reference: untracked Isend + blocking receive + later barrier; here stream-ordered)
static inline void comm_asynchSendArray(DeviceAmpArray& toSend, Index numAmpsToSend, Nat pairRank) {
    DFSA_CHECK(dfsa_x_send(toSend.owner, toSend.which, 0, toSend.which, 0, numAmpsToSend, int(pairRank)));
}
static inline void comm_receiveArray(DeviceAmpArray& toReceive, Index numAmpsToReceive, Nat pairRank) {
    DFSA_CHECK(dfsa_x_recv(toReceive.owner, toReceive.which, 0, numAmpsToReceive, int(pairRank)));
}

static inline void comm_reduceAmp(Amp& localAmp) {
    double reim[2] = {localAmp.real(), localAmp.imag()};
    DFSA_CHECK(dfsa_x_allreduce_amp(reim));
    localAmp = Amp(reim[0], reim[1]);
}
