// dfsa_comm.cu -- process bootstrap and the pairwise amplitude exchange of libdfsa_b200.so.
//
// Replaces src/communication.hpp of the reference (MPI_Init/Barrier/Isend/Irecv/Waitall/Allreduce). One process
// per GPU. Two transports behind the same dfsa_x_* calls:
//   "nccl": ncclSend + ncclRecv in one group on the comm stream (NVLink 5 / NVSwitch between the GPUs of one box);
//           stream-ordered against the compute stream with events -- no host synchronisation on the data path.
//   "ipc" : cudaIpc-mapped peer shards + cudaMemcpyAsync puts, host-mediated pair barriers in shared memory.
//           Works for several ranks on ONE device (NCCL refuses that), which is what lets the multi-rank parity
//           tests run on a single-GPU box; it is also the groundwork for the fused remote-load kernels.
// Host side channel: one POSIX shared-memory page set (barrier, NCCL id, IPC handle registry, reduce slots).
#include <errno.h>
#include <stddef.h>
#include <fcntl.h>
#include <map>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>
#include <vector>

#include "dfsa_internal.cuh"

namespace {

constexpr int      MAXP = 16;
constexpr int      MAXALLOC = 128;
constexpr uint32_t SHM_MAGIC = 0xDF5AB200u;

struct AllocSlot {
    cudaIpcMemHandle_t handle;
    uint64_t           bytes;
    volatile int       valid;
};

struct Shm {
    volatile uint32_t magic;
    int               size;
    volatile int      barrierCount, barrierSense;
    volatile int      idReady;
    char              ncclId[128];
    double            reduce[MAXP][4];
    volatile uint64_t pairSeq[MAXP][MAXP];
    AllocSlot         alloc[MAXP][MAXALLOC];
    volatile int      curSlot[MAXP][MAXALLOC][2];   // [rank][state key][amps|buffer] -> registry slot currently playing that role
    // stream-ordered signalling between the ranks' compute streams (no host synchronisation around a fused exchange):
    volatile uint64_t progress[MAXP];               // written BY THE DEVICE (stream memory op): rank r's compute stream has reached ticket progress[r]
    volatile uint64_t ticket[MAXP][MAXP][2];        // written by the host: ticket[r][p][n & 1] = r's "ready" ticket of its n-th rendezvous with p
    volatile int      ticketSlot[MAXP][MAXP][2];    // ... and the registry slot of the array r offers for reading in that rendezvous
    volatile int      device[MAXP];                 // CUDA device of each rank
    volatile int      peerOk[MAXP];                 // rank r can map every other rank's shards (peer access / same device)
    volatile int      signalOk[MAXP];               // rank r has stream memory ops on the shared page
    // expecPauliString: each rank's partial sum lands here straight from its reduction kernel, flag = sequence number
    volatile double   expecVal[2][MAXP][2];
    volatile uint64_t expecFlag[2][MAXP];
};

struct CommState {
    Shm*        shm = nullptr;
    bool        shmOwner = false;
    std::string shmName;
    int         localSense = 0;
    ncclComm_t  nccl = nullptr;
    bool        forked = false;
    std::vector<pid_t> children;
    bool        slotUsed[MAXALLOC] = {};
    void*       localPtr[MAXALLOC] = {};
    std::map<std::pair<int, int>, void*> peerMap;   // (rank, slot) -> mapped device pointer
    uint64_t    pairCount[MAXP] = {};
    // stream-ordered signalling
    bool        signals = false;                    // every rank can wait on / write to the shared page from its stream
    bool        peersOk = false;                    // every rank can map every other rank's shards
    int         fusedOverride = -1;                 // dfsa_comm_set_fused: -1 = environment decides
    void*       shmDev = nullptr;                   // device address of the shared page (cudaHostRegister)
    bool        shmRegistered = false;
    cudaEvent_t evXStart = nullptr, evXStop = nullptr;   // around the kernel(s) of the last fused exchange step (measurement)
    uint64_t    myTicket = 0;                       // last ticket this rank posted on its compute stream
    uint64_t    expecSeq = 0;
};

CommState g_comm;

#define DFSA_NCCL(call)                                                                              \
    do {                                                                                             \
        ncclResult_t r_ = (call);                                                                    \
        if (r_ != ncclSuccess) {                                                                     \
            dfsaSetError("%s:%d: %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(r_));      \
            return DFSA_ERR_NCCL;                                                                    \
        }                                                                                            \
    } while (0)

void napBriefly() {
    struct timespec ts = {0, 50000};
    nanosleep(&ts, nullptr);
}

// how long a rank waits for its peers before declaring the job dead (DFSA_COMM_TIMEOUT_S, default 120 s)
double commTimeoutSeconds() {
    static double t = -1;
    if (t < 0) { const char* e = getenv("DFSA_COMM_TIMEOUT_S"); t = e ? atof(e) : 120.0; if (t <= 0) t = 120.0; }
    return t;
}
double nowSeconds() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
#define DFSA_TRACE(...) do { if (getenv("DFSA_TRACE")) { fprintf(stderr, "[dfsa %d] ", dfsaCtx().rank); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); fflush(stderr); } } while (0)

int shmBarrier() {
    Shm* m = g_comm.shm;
    if (!m || m->size == 1) return DFSA_OK;
    g_comm.localSense = !g_comm.localSense;
    if (__sync_add_and_fetch(&m->barrierCount, 1) == m->size) {
        m->barrierCount = 0;
        __sync_synchronize();
        m->barrierSense = g_comm.localSense;
    } else {
        uint64_t spins = 0;
        const double t0 = nowSeconds();
        while (m->barrierSense != g_comm.localSense) {
            if (++spins > 2000) {
                napBriefly();
                if ((spins & 1023) == 0 && nowSeconds() - t0 > commTimeoutSeconds()) { dfsaSetError("host barrier timed out (a rank died?)"); return DFSA_ERR_COMM; }
            }
        }
    }
    __sync_synchronize();
    return DFSA_OK;
}

// both partners call this the same number of times; returns once the partner has arrived too
int pairBarrier(int pair) {
    Shm* m = g_comm.shm;
    DfsaContext& c = dfsaCtx();
    uint64_t mine = ++g_comm.pairCount[pair];
    __sync_synchronize();
    m->pairSeq[c.rank][pair] = mine;
    uint64_t spins = 0;
    const double t0 = nowSeconds();
    while (m->pairSeq[pair][c.rank] < mine) {
        if (++spins > 2000) {
            napBriefly();
            if ((spins & 1023) == 0 && nowSeconds() - t0 > commTimeoutSeconds()) { dfsaSetError("pair barrier with rank %d timed out", pair); return DFSA_ERR_COMM; }
        }
    }
    __sync_synchronize();
    return DFSA_OK;
}

size_t shmBytes() {
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    return (sizeof(Shm) + page - 1) / page * page;
}

int mapShm(const char* name, bool create, int size) {
    size_t bytes = shmBytes();
    int fd = -1;
    if (create) {
        shm_unlink(name);
        fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0) { dfsaSetError("shm_open/ftruncate(%s) failed: %s", name, strerror(errno)); return DFSA_ERR_COMM; }
    } else {
        for (int tries = 0; tries < 600000; tries++) {       // up to ~60 s
            fd = shm_open(name, O_RDWR, 0600);
            if (fd >= 0) {
                struct stat st;
                if (fstat(fd, &st) == 0 && (size_t)st.st_size >= bytes) break;
                close(fd); fd = -1;
            }
            struct timespec ts = {0, 100000};
            nanosleep(&ts, nullptr);
        }
        if (fd < 0) { dfsaSetError("rank 0 never created the shared segment %s", name); return DFSA_ERR_COMM; }
    }
    void* mem = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (mem == MAP_FAILED) { dfsaSetError("mmap of %s failed: %s", name, strerror(errno)); return DFSA_ERR_COMM; }
    g_comm.shm = (Shm*)mem;
    if (create) {
        memset(mem, 0, bytes);
        g_comm.shm->size = size;
        __sync_synchronize();
        g_comm.shm->magic = SHM_MAGIC;
    } else {
        for (uint64_t spins = 0; g_comm.shm->magic != SHM_MAGIC; spins++) {
            napBriefly();
            if (spins > 1200000) { dfsaSetError("shared segment %s never became ready", name); return DFSA_ERR_COMM; }
        }
    }
    return DFSA_OK;
}

// ---- stream memory operations (driver API through the runtime's entry-point lookup: no link dependency on libcuda)
typedef int (*StreamValueFn)(cudaStream_t, unsigned long long /*CUdeviceptr*/, unsigned long long /*value*/, unsigned int /*flags*/);
StreamValueFn g_streamWrite64 = nullptr, g_streamWait64 = nullptr;

bool loadStreamMemOps() {
    if (g_streamWrite64 && g_streamWait64) return true;
    void *w = nullptr, *q = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue64", &w, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) return false;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue64", &q, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) return false;
    g_streamWrite64 = (StreamValueFn)w;
    g_streamWait64 = (StreamValueFn)q;
    return w && q;
}

inline unsigned long long progressDevAddr(int rank) {
    return (unsigned long long)((char*)g_comm.shmDev + offsetof(Shm, progress) + sizeof(uint64_t) * (size_t)rank);
}

// "everything enqueued on my compute stream so far is done" becomes visible to the other ranks as progress[me] >= ticket
int postTicket(uint64_t* ticketOut) {
    DfsaContext& c = dfsaCtx();
    const uint64_t t = ++g_comm.myTicket;
    if (g_streamWrite64(c.compute, progressDevAddr(c.rank), t, 0 /*CU_STREAM_WRITE_VALUE_DEFAULT: fenced*/) != 0) {
        dfsaSetError("cuStreamWriteValue64 failed");
        return DFSA_ERR_CUDA;
    }
    *ticketOut = t;
    return DFSA_OK;
}

// my compute stream goes no further until rank `peer`'s stream has reached `ticket`
int awaitTicket(int peer, uint64_t ticket) {
    if (g_streamWait64(dfsaCtx().compute, progressDevAddr(peer), ticket, 0 /*CU_STREAM_WAIT_VALUE_GEQ*/) != 0) {
        dfsaSetError("cuStreamWaitValue64 failed");
        return DFSA_ERR_CUDA;
    }
    return DFSA_OK;
}

// Host rendezvous with a set of peers (no stream is touched): publish my ticket to each, then collect theirs. Everybody posts
// before anybody waits, so any set of mutually consistent groups is deadlock-free. Slots are double-buffered by the parity
// of the pair's rendezvous count: a slot is rewritten two rendezvous later, which the peer only lets happen after reading it.
int exchangeTickets(const int* peers, int n, uint64_t mine, int mySlot, uint64_t* theirs, int* theirSlots) {
    Shm* m = g_comm.shm;
    DfsaContext& c = dfsaCtx();
    uint64_t count[MAXP];
    for (int i = 0; i < n; i++) {
        const int p = peers[i];
        count[i] = ++g_comm.pairCount[p];
        m->ticket[c.rank][p][count[i] & 1] = mine;
        m->ticketSlot[c.rank][p][count[i] & 1] = mySlot;
        __sync_synchronize();
        m->pairSeq[c.rank][p] = count[i];
    }
    const double t0 = nowSeconds();
    for (int i = 0; i < n; i++) {
        const int p = peers[i];
        uint64_t spins = 0;
        while (m->pairSeq[p][c.rank] < count[i]) {
            if (++spins > 2000) {
                napBriefly();
                if ((spins & 1023) == 0 && nowSeconds() - t0 > commTimeoutSeconds()) { dfsaSetError("rendezvous with rank %d timed out", p); return DFSA_ERR_COMM; }
            }
        }
        __sync_synchronize();
        theirs[i] = m->ticket[p][c.rank][count[i] & 1];
        theirSlots[i] = m->ticketSlot[p][c.rank][count[i] & 1];
    }
    return DFSA_OK;
}

int finishInit() {
    DfsaContext& c = dfsaCtx();
    c.initialised = true;
    if (c.size == 1) { c.transport = Transport::Single; return dfsaEnsureDevice(); }

    int devCount = 0;
    cudaError_t e = cudaGetDeviceCount(&devCount);
    if (e != cudaSuccess || devCount == 0) { dfsaSetError("no CUDA device available; libdfsa_b200 has no CPU fallback"); return DFSA_ERR_CUDA; }
    const char* lr = getenv("LOCAL_RANK");
    int local = lr ? atoi(lr) : c.rank;
    c.device = local % devCount;
    DFSA_TRY(dfsaEnsureDevice());

    const char* want = getenv("DFSA_TRANSPORT");
    bool distinctDevices = devCount >= c.size;
    bool useNccl = want ? (strcmp(want, "nccl") == 0) : distinctDevices;
    if (useNccl && !distinctDevices) { dfsaSetError("DFSA_TRANSPORT=nccl needs one GPU per rank (%d ranks, %d devices)", c.size, devCount); return DFSA_ERR_COMM; }

    if (useNccl) {
        ncclUniqueId id;
        static_assert(sizeof(ncclUniqueId) == 128, "unique id size");
        Shm* m = g_comm.shm;
        if (c.rank == 0) {
            DFSA_NCCL(ncclGetUniqueId(&id));
            memcpy(m->ncclId, &id, sizeof(id));
            __sync_synchronize();
            m->idReady = 1;
        } else {
            for (uint64_t spins = 0; !m->idReady; spins++) {
                napBriefly();
                if (spins > 1200000) { dfsaSetError("NCCL id never published"); return DFSA_ERR_COMM; }
            }
            __sync_synchronize();
            memcpy(&id, m->ncclId, sizeof(id));
        }
        DFSA_NCCL(ncclCommInitRank(&g_comm.nccl, c.size, id, c.rank));
        c.transport = Transport::Nccl;
    } else {
        c.transport = Transport::Ipc;
    }
    // Capabilities every rank must have for the fused remote-load kernels, agreed collectively so that all ranks take the same
    // path: (a) this rank can map every other rank's shards (same device, or peer access between the two GPUs), (b) stream
    // memory operations on the shared page (stream-ordered signalling; without them the fused kernels are host-synchronised).
    Shm* m = g_comm.shm;
    m->device[c.rank] = c.device;
    DFSA_TRY(shmBarrier());
    int ok = 1;
    for (int r = 0; r < c.size && ok; r++) {
        if (r == c.rank || m->device[r] == c.device) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, c.device, m->device[r]) != cudaSuccess || !can) ok = 0;
    }
    cudaGetLastError();
    m->peerOk[c.rank] = ok;
    int sig = 0;
    const char* noSig = getenv("DFSA_STREAM_SIGNALS");
    if (!(noSig && atoi(noSig) == 0) && loadStreamMemOps()) {
        const size_t page = (size_t)sysconf(_SC_PAGESIZE), bytes = (sizeof(Shm) + page - 1) / page * page;
        if (cudaHostRegister((void*)m, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable) == cudaSuccess) {
            g_comm.shmRegistered = true;
            if (cudaHostGetDevicePointer(&g_comm.shmDev, (void*)m, 0) == cudaSuccess && g_comm.shmDev) sig = 1;
        }
        cudaGetLastError();
    }
    m->signalOk[c.rank] = sig;
    DFSA_TRY(shmBarrier());
    g_comm.peersOk = g_comm.signals = true;
    for (int r = 0; r < c.size; r++) { g_comm.peersOk = g_comm.peersOk && m->peerOk[r]; g_comm.signals = g_comm.signals && m->signalOk[r]; }
    g_comm.myTicket = 0;
    return shmBarrier();
}

}  // namespace

// ------------------------------------------------------------------------------------------------ bootstrap

extern "C" int dfsa_comm_init(void) {
    DfsaContext& c = dfsaCtx();
    if (c.initialised) return DFSA_OK;                        // idempotent like comm_init (communication.hpp:16-21)
    const char* ws = getenv("WORLD_SIZE");
    const char* rk = getenv("RANK");
    const char* np = getenv("DFSA_NP");
    if (ws && rk && atoi(ws) > 1) {
        c.size = atoi(ws);
        c.rank = atoi(rk);
        if (c.size > MAXP) { dfsaSetError("at most %d ranks", MAXP); return DFSA_ERR_ARG; }
        // One node, at most MAXP ranks: the side channel is a POSIX shared-memory segment and the fused kernels read peer
        // shards over NVLink. A launcher that spreads the job over several hosts is refused here instead of hanging.
        const char* lws = getenv("LOCAL_WORLD_SIZE");
        if (lws && atoi(lws) != c.size) {
            dfsaSetError("WORLD_SIZE=%d but LOCAL_WORLD_SIZE=%s: libdfsa_b200 runs the ranks of ONE node (multi-node needs dfsa_comm_init_with_id and the staged NCCL transport)", c.size, lws);
            return DFSA_ERR_COMM;
        }
        // Segment name: DFSA_JOB_ID if given, else the rendezvous the launcher already agreed on (MASTER_ADDR:MASTER_PORT) --
        // identical on every rank whatever process spawned it (srun / mpirun wrappers give each rank its own parent).
        const char* job = getenv("DFSA_JOB_ID");
        const char* port = getenv("MASTER_PORT");
        const char* addr = getenv("MASTER_ADDR");
        char name[160];
        if (job) snprintf(name, sizeof(name), "/dfsa_%s", job);
        else if (port) {
            unsigned h = 2166136261u;
            for (const char* q = addr ? addr : ""; *q; q++) h = (h ^ (unsigned char)*q) * 16777619u;
            snprintf(name, sizeof(name), "/dfsa_%08x_%s_u%d", h, port, (int)getuid());
        } else {
            dfsaSetError("RANK/WORLD_SIZE are set but neither DFSA_JOB_ID nor MASTER_PORT: the ranks cannot agree on a shared segment");
            return DFSA_ERR_COMM;
        }
        g_comm.shmName = name;
        g_comm.shmOwner = (c.rank == 0);
        DFSA_TRY(mapShm(name, c.rank == 0, c.size));
    } else if (np && atoi(np) > 1) {
        // fork model: must happen before this process touches CUDA (children cannot inherit a CUDA context)
        if (c.compute) { dfsaSetError("dfsa_comm_init with DFSA_NP must be the first CUDA-touching call"); return DFSA_ERR_COMM; }
        c.size = atoi(np);
        if (c.size > MAXP) { dfsaSetError("at most %d ranks", MAXP); return DFSA_ERR_ARG; }
        void* mem = mmap(nullptr, shmBytes(), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
        if (mem == MAP_FAILED) { dfsaSetError("mmap failed: %s", strerror(errno)); return DFSA_ERR_COMM; }
        memset(mem, 0, shmBytes());
        g_comm.shm = (Shm*)mem;
        g_comm.shm->size = c.size;
        g_comm.shm->magic = SHM_MAGIC;
        g_comm.forked = true;
        fflush(stdout); fflush(stderr);
        for (int r = 1; r < c.size; r++) {
            pid_t pid = fork();
            if (pid < 0) { dfsaSetError("fork failed: %s", strerror(errno)); return DFSA_ERR_COMM; }
            if (pid == 0) {
                c.rank = r;
                g_comm.children.clear();
                if (!getenv("DFSA_KEEP_STDOUT")) { if (!freopen("/dev/null", "w", stdout)) {} }
                break;
            }
            g_comm.children.push_back(pid);
        }
    } else {
        c.size = 1;
        c.rank = 0;
    }
    return finishInit();
}

extern "C" int dfsa_comm_get_unique_id(void* out128) {
    DFSA_REQUIRE(out128, "null id buffer");
    ncclUniqueId id;
    DFSA_NCCL(ncclGetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return DFSA_OK;
}

extern "C" int dfsa_comm_init_with_id(int rank, int numRanks, const void* uniqueId128, int device) {
    DfsaContext& c = dfsaCtx();
    if (c.initialised) return DFSA_OK;
    DFSA_REQUIRE(numRanks >= 1 && rank >= 0 && rank < numRanks && numRanks <= MAXP, "bad rank / size");
    c.rank = rank;
    c.size = numRanks;
    c.initialised = true;
    int devCount = 0;
    cudaError_t e = cudaGetDeviceCount(&devCount);
    if (e != cudaSuccess || devCount == 0) { dfsaSetError("no CUDA device available; libdfsa_b200 has no CPU fallback"); return DFSA_ERR_CUDA; }
    c.device = (device >= 0 ? device : rank) % devCount;
    DFSA_TRY(dfsaEnsureDevice());
    if (numRanks == 1) { c.transport = Transport::Single; return DFSA_OK; }
    DFSA_REQUIRE(uniqueId128, "null unique id");
    ncclUniqueId id;
    memcpy(&id, uniqueId128, sizeof(id));
    DFSA_NCCL(ncclCommInitRank(&g_comm.nccl, numRanks, id, rank));
    c.transport = Transport::Nccl;
    return DFSA_OK;
}

int dfsaHostBarrier() {
    DfsaContext& c = dfsaCtx();
    if (c.size == 1) return DFSA_OK;
    if (g_comm.shm) return shmBarrier();
    // NCCL-only bootstrap (no shared segment): a 1-element all-reduce is the barrier
    double2* scratch;
    DFSA_TRY(dfsaScratch(64, &scratch));
    DFSA_NCCL(ncclAllReduce(scratch, scratch, 1, ncclDouble, ncclSum, g_comm.nccl, c.comm));
    DFSA_CUDA(cudaStreamSynchronize(c.comm));
    return DFSA_OK;
}

extern "C" int dfsa_comm_barrier(void) {
    DFSA_TRY(dfsa_device_sync());
    return dfsaHostBarrier();
}

extern "C" int dfsa_comm_finalize(void) {
    DfsaContext& c = dfsaCtx();
    if (!c.initialised) return DFSA_OK;
    DFSA_TRY(dfsa_comm_barrier());                            // comm_end: Barrier then Finalize (communication.hpp:24-27)
    DFSA_TRY(dfsaPoolDrain());                                // recycled shards: unregister (collective) and free
    for (auto& kv : g_comm.peerMap) cudaIpcCloseMemHandle(kv.second);
    g_comm.peerMap.clear();
    // nobody may free a shard (states destroyed after comm_end skip the collective unregister) while a slower rank still
    // holds a mapping of it
    DFSA_TRY(dfsaHostBarrier());
    if (g_comm.nccl) { ncclCommDestroy(g_comm.nccl); g_comm.nccl = nullptr; }
    if (g_comm.shmRegistered) { cudaHostUnregister((void*)g_comm.shm); g_comm.shmRegistered = false; }
    if (g_comm.forked) {
        fflush(stdout);
        if (c.rank != 0) _exit(0);                            // children never run the caller's epilogue
        for (pid_t pid : g_comm.children) { int st; waitpid(pid, &st, 0); }
    }
    if (g_comm.shm) {
        if (g_comm.shmOwner) shm_unlink(g_comm.shmName.c_str());
        munmap((void*)g_comm.shm, shmBytes());
    }
    // a later comm_init starts from a clean slate (fresh segment, fresh counters)
    g_comm = CommState();
    c.initialised = false;
    c.size = 1; c.rank = 0; c.transport = Transport::Single;
    return DFSA_OK;
}

extern "C" int dfsa_comm_rank(void) { return dfsaCtx().rank; }
extern "C" int dfsa_comm_size(void) { return dfsaCtx().size; }
extern "C" const char* dfsa_comm_transport(void) {
    switch (dfsaCtx().transport) { case Transport::Nccl: return "nccl"; case Transport::Ipc: return "ipc"; default: return "single"; }
}

// ------------------------------------------------------------------------------------------------ IPC registry

int dfsaRegisterAllocation(void* ptr, size_t bytes, int* idOut) {
    DfsaContext& c = dfsaCtx();
    int slot = -1;
    for (int i = 0; i < MAXALLOC; i++) if (!g_comm.slotUsed[i]) { slot = i; break; }   // same sequence on every rank (SPMD)
    if (slot < 0) { dfsaSetError("too many live states (%d allocation slots)", MAXALLOC); return DFSA_ERR_UNSUPPORTED; }
    g_comm.slotUsed[slot] = true;
    g_comm.localPtr[slot] = ptr;
    *idOut = slot;
    if (!g_comm.shm) return DFSA_OK;                           // id-only bootstrap: no side channel, no peer mapping
    AllocSlot& a = g_comm.shm->alloc[c.rank][slot];
    DFSA_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)&a.handle, ptr));
    a.bytes = bytes;
    __sync_synchronize();
    a.valid = 1;
    // No barrier: a peer only ever looks a handle up inside an exchange, after the pair (or global) barrier that opens it,
    // and this rank reaches that barrier after this line in program order. State creation stays a local operation,
    // like the reference's constructor (states.hpp:31-48).
    __sync_synchronize();
    return DFSA_OK;
}

int dfsaPublishArrays(dfsa_state* s) {
    if (!g_comm.shm || s->key < 0) return DFSA_OK;
    DfsaContext& c = dfsaCtx();
    g_comm.shm->curSlot[c.rank][s->key][0] = s->allocId[0];
    g_comm.shm->curSlot[c.rank][s->key][1] = s->allocId[1];
    __sync_synchronize();
    return DFSA_OK;
}

int dfsaUnregisterAllocation(int id) {
    DfsaContext& c = dfsaCtx();
    if (g_comm.shm) {
        DFSA_TRY(shmBarrier());                                // nobody is still using the mappings
        for (auto it = g_comm.peerMap.begin(); it != g_comm.peerMap.end();) {
            if (it->first.second == id) { cudaIpcCloseMemHandle(it->second); it = g_comm.peerMap.erase(it); }
            else ++it;
        }
        g_comm.shm->alloc[c.rank][id].valid = 0;
        DFSA_TRY(shmBarrier());                                // all mappings closed before the owner frees
    }
    g_comm.slotUsed[id] = false;
    g_comm.localPtr[id] = nullptr;
    return DFSA_OK;
}

static int peerPointer(int pair, int slot, double2** out) {
    auto key = std::make_pair(pair, slot);
    auto it = g_comm.peerMap.find(key);
    if (it == g_comm.peerMap.end()) {
        AllocSlot& a = g_comm.shm->alloc[pair][slot];
        if (!a.valid) { dfsaSetError("rank %d has not published allocation %d", pair, slot); return DFSA_ERR_COMM; }
        void* p = nullptr;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const void*)&a.handle, sizeof(h));
        DFSA_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        it = g_comm.peerMap.emplace(key, p).first;
    }
    *out = (double2*)it->second;
    return DFSA_OK;
}

// device pointer (mapped into this process) of rank `pair`'s current amps / buffer array of state `s`
static int peerArray(dfsa_state* s, int pair, int which, double2** out) {
    if (!g_comm.shm || s->key < 0) { dfsaSetError("no shared side channel: peer arrays are not mapped"); return DFSA_ERR_COMM; }
    __sync_synchronize();
    return peerPointer(pair, g_comm.shm->curSlot[pair][s->key][which], out);
}

// ------------------------------------------------------------------------------------------------ exchange

static int checkXArgs(dfsa_state* s, int sendWhich, uint64_t sendStart, int recvWhich, uint64_t recvStart, uint64_t num, int pairRank, bool bothWays = true) {
    DfsaContext& c = dfsaCtx();
    DFSA_REQUIRE(s && c.size > 1, "exchange needs more than one rank");
    DFSA_REQUIRE(pairRank >= 0 && pairRank < c.size && pairRank != c.rank, "bad pair rank");
    DFSA_REQUIRE((sendWhich | 1) == 1 && (recvWhich | 1) == 1, "array selector must be DFSA_AMPS or DFSA_BUFFER");
    DFSA_REQUIRE(sendStart + num <= s->numAmps && recvStart + num <= s->numAmps, "exchange range exceeds the shard");
    // a rank that both sends and receives must not receive into what it is still sending (the MPI no-overlap rule)
    DFSA_REQUIRE(!bothWays || sendWhich != recvWhich || sendStart + num <= recvStart || recvStart + num <= sendStart, "send and receive regions overlap");
    return DFSA_OK;
}

// direction flags: doSend / doRecv
static int transfer(dfsa_state* s, int sendWhich, uint64_t sendStart, int recvWhich, uint64_t recvStart, uint64_t num, int pairRank, bool doSend, bool doRecv) {
    DfsaContext& c = dfsaCtx();
    if (c.transport == Transport::Nccl) {
        // comm stream waits for every kernel enqueued so far; compute stream then waits for the transfer
        DFSA_CUDA(cudaEventRecord(c.evCompute, c.compute));
        DFSA_CUDA(cudaStreamWaitEvent(c.comm, c.evCompute, 0));
        DFSA_NCCL(ncclGroupStart());
        if (doSend) DFSA_NCCL(ncclSend(s->arr[sendWhich] + sendStart, 2 * num, ncclDouble, pairRank, g_comm.nccl, c.comm));
        if (doRecv) DFSA_NCCL(ncclRecv(s->arr[recvWhich] + recvStart, 2 * num, ncclDouble, pairRank, g_comm.nccl, c.comm));
        DFSA_NCCL(ncclGroupEnd());
        DFSA_CUDA(cudaEventRecord(c.evComm, c.comm));
        DFSA_CUDA(cudaStreamWaitEvent(c.compute, c.evComm, 0));
        return DFSA_OK;
    }
    // IPC: put into the partner's receive region
    DFSA_TRACE("ipc transfer pair=%d num=%llu send=%d recv=%d", pairRank, (unsigned long long)num, (int)doSend, (int)doRecv);
    DFSA_CUDA(cudaStreamSynchronize(c.compute));               // my data is final, my receive region is free
    DFSA_TRY(pairBarrier(pairRank));                           // ... and so is the partner's
    if (doSend) {
        double2* remote;
        DFSA_TRY(peerArray(s, pairRank, recvWhich, &remote));
        DFSA_CUDA(cudaMemcpyAsync(remote + recvStart, s->arr[sendWhich] + sendStart, num * sizeof(double2), cudaMemcpyDeviceToDevice, c.comm));
        DFSA_CUDA(cudaStreamSynchronize(c.comm));
    }
    return pairBarrier(pairRank);                              // the partner's put has landed in my region
}

extern "C" int dfsa_x_exchange(dfsa_state* s, int sendWhich, uint64_t sendStart, int recvWhich, uint64_t recvStart, uint64_t num, int pairRank) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_TRY(checkXArgs(s, sendWhich, sendStart, recvWhich, recvStart, num, pairRank));
    return transfer(s, sendWhich, sendStart, recvWhich, recvStart, num, pairRank, true, true);
}

extern "C" int dfsa_x_send(dfsa_state* s, int sendWhich, uint64_t sendStart, int recvWhich, uint64_t recvStart, uint64_t num, int pairRank) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_TRY(checkXArgs(s, sendWhich, sendStart, recvWhich, recvStart, num, pairRank, false));
    return transfer(s, sendWhich, sendStart, recvWhich, recvStart, num, pairRank, true, false);
}

extern "C" int dfsa_x_recv(dfsa_state* s, int recvWhich, uint64_t recvStart, uint64_t num, int pairRank) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_TRY(checkXArgs(s, recvWhich, recvStart, recvWhich, recvStart, num, pairRank, false));
    return transfer(s, recvWhich, recvStart, recvWhich, recvStart, num, pairRank, false, true);
}

// ------------------------------------------------------------------------------------------------ pipelined exchange + combine

namespace {
constexpr int kMaxChunks = 16;
cudaEvent_t g_chunkEvents[kMaxChunks];
bool g_chunkEventsReady = false;

// how many pieces the shard travels in: >= 64 MiB each, at most kMaxChunks
int chunkCountFor(uint64_t numAmps) {
    const char* e = getenv("DFSA_XCHG_CHUNKS");
    int want = e ? atoi(e) : kMaxChunks;
    if (want < 1) want = 1;
    if (want > kMaxChunks) want = kMaxChunks;
    const char* m = getenv("DFSA_XCHG_MIN_CHUNK_LOG2");          // smallest chunk, log2 amplitudes (default 2^22 = 64 MiB)
    const unsigned minLog = m ? (unsigned)atoi(m) : 22u;
    while (want > 1 && (numAmps / want) < (1ULL << minLog)) want >>= 1;
    int c = 1;
    while (c * 2 <= want) c *= 2;          // power of two so chunks tile the (power-of-two) shard exactly
    return c;
}

template <class Combine>
int pipelinedExchange(dfsa_state* s, int pairRank, int chunks, Combine combineRange) {
    DfsaContext& c = dfsaCtx();
    if (!g_chunkEventsReady) {
        for (int i = 0; i < kMaxChunks; i++) DFSA_CUDA(cudaEventCreateWithFlags(&g_chunkEvents[i], cudaEventDisableTiming));
        g_chunkEventsReady = true;
    }
    const uint64_t per = s->numAmps / chunks;
    DFSA_CUDA(cudaEventRecord(c.evCompute, c.compute));
    DFSA_CUDA(cudaStreamWaitEvent(c.comm, c.evCompute, 0));
    for (int k = 0; k < chunks; k++) {
        const uint64_t first = per * k;
        DFSA_NCCL(ncclGroupStart());
        DFSA_NCCL(ncclSend(s->arr[DFSA_AMPS] + first, 2 * per, ncclDouble, pairRank, g_comm.nccl, c.comm));
        DFSA_NCCL(ncclRecv(s->arr[DFSA_BUFFER] + first, 2 * per, ncclDouble, pairRank, g_comm.nccl, c.comm));
        DFSA_NCCL(ncclGroupEnd());
        DFSA_CUDA(cudaEventRecord(g_chunkEvents[k], c.comm));
        // chunk k of amps has left and chunk k of the partner has arrived: combine it while chunk k+1 is in flight
        DFSA_CUDA(cudaStreamWaitEvent(c.compute, g_chunkEvents[k], 0));
        DFSA_TRY(combineRange(first, per));
    }
    return DFSA_OK;
}
}  // namespace

double2 dfsaPowIHost(unsigned k);

// Fused remote-load form: usable when the ranks share a node-local side channel (peer shards mapped through CUDA IPC).
// Host protocol per op: [my earlier kernels done] pair barrier [launch: read partner amps over NVLink, write my buffer]
// [kernel done] pair barrier [swap amps<->buffer]. The second barrier keeps a partner from overwriting the shard this
// rank is still reading. DFSA_FUSED_EXCHANGE=0 falls back to the staged NCCL / copy-engine path.
static bool fusedAvailable() {
    static int env = -1;
    if (env < 0) { const char* e = getenv("DFSA_FUSED_EXCHANGE"); env = (e && atoi(e) == 0) ? 0 : 1; }
    const int on = g_comm.fusedOverride >= 0 ? g_comm.fusedOverride : env;
    // the side channel exists (ranks share a node) AND every rank can map every other rank's shards; otherwise the staged
    // NCCL / copy-engine path is used -- decided collectively at init, so all ranks agree
    return on == 1 && g_comm.shm != nullptr && g_comm.peersOk;
}

// Collective: switch between the fused remote-load kernels (1), the staged pack / exchange / combine path (0), or back to
// what DFSA_FUSED_EXCHANGE says (-1). bench.py's self-check runs the same circuit both ways.
extern "C" int dfsa_comm_set_fused(int mode) {
    DFSA_REQUIRE(mode >= -1 && mode <= 1, "mode is -1, 0 or 1");
    DFSA_TRY(dfsa_comm_barrier());
    g_comm.fusedOverride = mode;
    return DFSA_OK;
}
// device time of the kernel of the most recent fused exchange step on this rank (between the peers' READY and this rank's
// DONE, so waiting for a late partner is not in it). Synchronises on that kernel. -1 if there was none.
extern "C" int dfsa_comm_last_exchange_ms(double* ms) {
    DFSA_REQUIRE(ms, "null argument");
    *ms = -1.0;
    if (!g_comm.evXStop) return DFSA_OK;
    float f = 0.f;
    DFSA_CUDA(cudaEventSynchronize(g_comm.evXStop));
    DFSA_CUDA(cudaEventElapsedTime(&f, g_comm.evXStart, g_comm.evXStop));
    *ms = f;
    return DFSA_OK;
}
extern "C" int dfsa_comm_fused_active(void) { return fusedAvailable() ? (g_comm.signals ? 2 : 1) : 0; }

// One fused step with the ranks in `peers` (the partner of a pairwise op; the 2^k - 1 other members of a relocation or
// quad-depolarising group): launch(remote[i] = peers[i]'s current amplitude array) must write this rank's `buffer`;
// afterwards amps <-> buffer. Ordering, all on the compute streams (stream memory operations on the shared page):
//   [my earlier kernels] -> post READY -> wait every peer's READY -> kernel -> post DONE -> wait every peer's DONE
// so a peer's shard is final before it is read and is not overwritten (it becomes that peer's buffer) while it is being read.
// The hosts only meet to trade ticket numbers and array slots (a few shared-memory words), never to drain a stream, so
// consecutive gates pipeline. Without stream memory ops: the same steps with host synchronisation (round-1 protocol).
static int peerPointer(int pair, int slot, double2** out);

static int swapArraysAfter(dfsa_state* s) { return dfsa_state_swap_arrays(s); }

template <class Launch, class After>
static int fusedGroupExchange(dfsa_state* s, const int* peers, int n, Launch launch, After after) {
    DfsaContext& c = dfsaCtx();
    const double2* remote[MAXP];
    if (g_comm.signals) {
        // transfers of the staged path run on the comm stream; the compute stream is already ordered after them (transfer())
        // (the array a peer offers travels with its ticket: by the time this rank looks, a faster peer may already have
        // finished the step and swapped its arrays, so the registry's "current" slot would be the wrong one)
        uint64_t ready = 0, theirs[MAXP];
        int theirSlots[MAXP];
        DFSA_TRY(postTicket(&ready));
        DFSA_TRY(exchangeTickets(peers, n, ready, s->allocId[DFSA_AMPS], theirs, theirSlots));
        for (int i = 0; i < n; i++) DFSA_TRY(awaitTicket(peers[i], theirs[i]));
        for (int i = 0; i < n; i++) { double2* p; DFSA_TRY(peerPointer(peers[i], theirSlots[i], &p)); remote[i] = p; }
        if (!g_comm.evXStart) { DFSA_CUDA(cudaEventCreate(&g_comm.evXStart)); DFSA_CUDA(cudaEventCreate(&g_comm.evXStop)); }
        DFSA_CUDA(cudaEventRecord(g_comm.evXStart, c.compute));      // after the waits: the kernel's own time, without rank skew
        DFSA_TRY(launch(remote));
        DFSA_CUDA(cudaEventRecord(g_comm.evXStop, c.compute));
        uint64_t done = 0;
        DFSA_TRY(postTicket(&done));                               // == ready + 1 on every rank: nothing else posts in between
        for (int i = 0; i < n; i++) DFSA_TRY(awaitTicket(peers[i], theirs[i] + 1));
        return after(s);
    }
    DFSA_CUDA(cudaStreamSynchronize(c.comm));
    DFSA_CUDA(cudaStreamSynchronize(c.compute));
    for (int i = 0; i < n; i++) DFSA_TRY(pairBarrier(peers[i]));
    for (int i = 0; i < n; i++) { double2* p; DFSA_TRY(peerArray(s, peers[i], DFSA_AMPS, &p)); remote[i] = p; }
    DFSA_TRY(launch(remote));
    DFSA_CUDA(cudaStreamSynchronize(c.compute));
    for (int i = 0; i < n; i++) DFSA_TRY(pairBarrier(peers[i]));
    return after(s);
}

template <class Launch>
static int fusedGroupExchange(dfsa_state* s, const int* peers, int n, Launch launch) { return fusedGroupExchange(s, peers, n, launch, swapArraysAfter); }

template <class Launch>
static int fusedExchange(dfsa_state* s, int pairRank, Launch launch) {
    return fusedGroupExchange(s, &pairRank, 1, [&](const double2* const* remote) { return launch(remote[0]); });
}

extern "C" int dfsa_xk_exchangeCombine(dfsa_state* s, int pairRank, const double f0[2], const double f1[2]) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(f0 && f1, "null factor");
    DFSA_TRY(checkXArgs(s, DFSA_AMPS, 0, DFSA_BUFFER, 0, s ? s->numAmps : 0, pairRank));
    const double2 c0 = make_double2(f0[0], f0[1]), c1 = make_double2(f1[0], f1[1]);
    const bool partnerFirst = dfsaCtx().rank > pairRank;          // this rank holds bit 1 of the target: same FMA nesting as the local kernel's upper output
    if (fusedAvailable())
        return fusedExchange(s, pairRank, [&](const double2* remote) { return dfsaLaunchFusedCombine(s, remote, c0, c1, partnerFirst); });
    const int chunks = (dfsaCtx().transport == Transport::Nccl) ? chunkCountFor(s->numAmps) : 1;
    if (chunks == 1) {
        DFSA_TRY(transfer(s, DFSA_AMPS, 0, DFSA_BUFFER, 0, s->numAmps, pairRank, true, true));
        return dfsaLaunchCombineRange(s, 0, s->numAmps, c0, c1, partnerFirst);
    }
    return pipelinedExchange(s, pairRank, chunks, [&](uint64_t first, uint64_t num) { return dfsaLaunchCombineRange(s, first, num, c0, c1, partnerFirst); });
}

extern "C" int dfsa_xk_swapSuffixPrefix(dfsa_state* s, unsigned qb1, unsigned movingBit, int pairRank) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s && qb1 < s->logNumAmps, "qb1 must be a suffix qubit");
    const uint64_t half = s->numAmps >> 1;
    DFSA_TRY(checkXArgs(s, DFSA_BUFFER, 0, DFSA_BUFFER, half, half, pairRank));
    movingBit &= 1u;
    if (fusedAvailable())                                                  // one pass: keep my half, read the partner's over NVLink
        return fusedExchange(s, pairRank, [&](const double2* remote) { return dfsaLaunchFusedSwap(s, remote, qb1, movingBit ^ 1u); });
    if (qb1 + 1 == s->logNumAmps) {
        // top suffix qubit: the moving half is contiguous, no packing (distributed_statevector.hpp:140-157)
        const uint64_t offset = half * movingBit;
        DFSA_TRY(transfer(s, DFSA_AMPS, offset, DFSA_BUFFER, 0, half, pairRank, true, true));
        return dfsa_k_copyFromBuffer(s, offset, 0, half);
    }
    const uint32_t pos = qb1;
    DFSA_TRY(dfsa_k_pack(s, &pos, 1, movingBit, 0));                       // buffer[0..A/2) = the half that leaves
    DFSA_TRY(transfer(s, DFSA_BUFFER, 0, DFSA_BUFFER, half, half, pairRank, true, true));
    return dfsa_k_unpack(s, &pos, 1, movingBit, half);
}

// Host-only part of the single-shot relocation: for this rank, rho = its bits of the swapped prefix qubits (bit i <-> pair i)
// and, for every value sigma of the k landing suffix bits, the rank whose shard supplies the amplitudes that end up at
// suffix bits sigma here: this rank with its swapped rank bits set to sigma. (After the swap, index bit s_i holds what
// was rank bit r_i and vice versa.)
extern "C" int dfsa_plan_relocate(int rank, unsigned logNumAmps, const uint32_t* prefixQubits, unsigned numPairs, int owners[16], unsigned* rhoOut) {
    DFSA_REQUIRE(prefixQubits && owners && rhoOut && numPairs >= 1 && numPairs <= 4 && rank >= 0, "bad argument");
    unsigned rho = 0;
    for (unsigned i = 0; i < numPairs; i++) {
        DFSA_REQUIRE(prefixQubits[i] >= logNumAmps, "not a prefix qubit");
        rho |= (((unsigned)rank >> (prefixQubits[i] - logNumAmps)) & 1u) << i;
    }
    for (unsigned sigma = 0; sigma < 16; sigma++) {
        int owner = rank;
        for (unsigned i = 0; i < numPairs; i++) {
            const unsigned rankBit = prefixQubits[i] - logNumAmps;
            owner = (owner & ~(1 << rankBit)) | (int)(((sigma >> i) & 1u) << rankBit);
        }
        owners[sigma] = sigma < (1u << numPairs) ? owner : -1;
    }
    *rhoOut = rho;
    return DFSA_OK;
}

// Relocation of manyTargGate (distributed_statevector.hpp:193-223): swap suffix qubit suffixQubits[i] with prefix qubit
// prefixQubits[i] for all i. COLLECTIVE over all ranks. One pair: the fused swap. Several pairs with peer-mapped shards: one
// pass that gathers from the 2^k shards of this rank's group ((1 - 2^-k) 16A bytes over NVLink instead of k * 8A, one
// pass over HBM instead of k). Otherwise the reference's sequence of swaps. Applying it twice restores the state.
extern "C" int dfsa_xk_relocate(dfsa_state* s, const uint32_t* suffixQubits, const uint32_t* prefixQubits, unsigned numPairs) {
    DFSA_TRY(dfsaEnsureDevice());
    DfsaContext& c = dfsaCtx();
    DFSA_REQUIRE(s && suffixQubits && prefixQubits && numPairs >= 1 && numPairs <= 4 && c.size > 1, "bad argument");
    const unsigned L = s->logNumAmps;
    for (unsigned i = 0; i < numPairs; i++) {
        DFSA_REQUIRE(suffixQubits[i] < L && prefixQubits[i] >= L && prefixQubits[i] < L + s->logNumNodes, "pairs are (suffix qubit, prefix qubit)");
        for (unsigned j = 0; j < i; j++)
            DFSA_REQUIRE(suffixQubits[i] != suffixQubits[j] && prefixQubits[i] != prefixQubits[j], "qubits must be distinct");
    }
    if (numPairs == 1 || !fusedAvailable()) {
        for (unsigned i = 0; i < numPairs; i++) {
            const unsigned rankBit = prefixQubits[i] - L, mine = ((unsigned)c.rank >> rankBit) & 1u;
            DFSA_TRY(dfsa_xk_swapSuffixPrefix(s, suffixQubits[i], mine ^ 1u, c.rank ^ (1 << rankBit)));
        }
        return DFSA_OK;
    }
    unsigned rho = 0;
    int owners[16];
    DFSA_TRY(dfsa_plan_relocate(c.rank, L, prefixQubits, numPairs, owners, &rho));
    int peers[16], slotOf[16], n = 0;
    for (unsigned sigma = 0; sigma < (1u << numPairs); sigma++) {
        slotOf[sigma] = -1;
        if (owners[sigma] != c.rank) { slotOf[sigma] = n; peers[n++] = owners[sigma]; }
    }
    return fusedGroupExchange(s, peers, n, [&](const double2* const* remote) {
        const double2* shards[16];
        for (unsigned sigma = 0; sigma < (1u << numPairs); sigma++) shards[sigma] = slotOf[sigma] < 0 ? s->arr[DFSA_AMPS] : remote[slotOf[sigma]];
        return dfsaLaunchRelocate(s, shards, suffixQubits, numPairs, rho);
    });
}

// oneQubitDepolarising on a qubit whose bra bit is a rank bit (distributed_densitymatrix.hpp:110-141)
extern "C" int dfsa_xk_depol1Prefix(dfsa_state* s, unsigned qb, unsigned bit, double prob, int pairRank) {
    DFSA_TRY(dfsaEnsureDevice());
    const uint64_t half = s ? s->numAmps >> 1 : 0;
    DFSA_TRY(checkXArgs(s, DFSA_BUFFER, 0, DFSA_BUFFER, half, half, pairRank));
    DFSA_REQUIRE(s->isDensity && qb < s->numQubits && qb >= s->numQubits - s->logNumNodes, "needs a prefix qubit of a density matrix");
    if (fusedAvailable())
        return fusedExchange(s, pairRank, [&](const double2* remote) { return dfsaLaunchFusedDepol1(s, remote, qb, bit, prob); });
    const uint32_t pos = qb;
    DFSA_TRY(dfsa_k_pack(s, &pos, 1, bit & 1u, 0));                        // the ket bit == bra bit half (reference :119-127)
    DFSA_TRY(transfer(s, DFSA_BUFFER, 0, DFSA_BUFFER, half, half, pairRank, true, true));
    return dfsa_k_depol1Combine(s, qb, bit, prob);
}

// damping on a qubit whose bra bit is a rank bit (distributed_densitymatrix.hpp:284-317): population flows one way, from
// the rank holding bra bit 1 to the rank holding bra bit 0
extern "C" int dfsa_xk_dampingPrefix(dfsa_state* s, unsigned qb, unsigned bit, double prob, int pairRank) {
    DFSA_TRY(dfsaEnsureDevice());
    const uint64_t half = s ? s->numAmps >> 1 : 0;
    DFSA_TRY(checkXArgs(s, DFSA_BUFFER, 0, DFSA_BUFFER, 0, half, pairRank, false));
    DFSA_REQUIRE(s->isDensity && qb < s->numQubits && qb >= s->numQubits - s->logNumNodes, "needs a prefix qubit of a density matrix");
    if (fusedAvailable())
        return fusedExchange(s, pairRank, [&](const double2* remote) { return dfsaLaunchFusedDamping(s, remote, qb, bit, prob); });
    if (bit & 1u) {
        DFSA_TRY(dfsa_k_dampingPrefix(s, qb, bit, prob, 0));
        DFSA_TRY(transfer(s, DFSA_BUFFER, 0, DFSA_BUFFER, 0, half, pairRank, true, false));
    }
    DFSA_TRY(dfsa_k_dampingPrefix(s, qb, bit, prob, 1));
    if (!(bit & 1u)) {
        DFSA_TRY(transfer(s, DFSA_BUFFER, 0, DFSA_BUFFER, 0, half, pairRank, false, true));
        DFSA_TRY(dfsa_k_dampingPrefix(s, qb, bit, prob, 2));
    }
    return DFSA_OK;     // the reference needs a global barrier here to protect the sender's buffer; stream order does that job
}

// Link microbenchmark (the NVLink line of the roofline must be measured on the box, not assumed): every rank pulls
// its partner's whole shard into its own exchange buffer, all pairs at once, both directions -- the traffic pattern of a
// prefix gate without the arithmetic. mode 0: remote loads from a kernel (what the fused kernels do), mode 1: copy engine
// (cudaMemcpyAsync from the peer mapping), mode 2: remote loads in ONE direction only (even ranks pull, odd ranks serve -- the
// traffic of the one-way damping transfer). *ms = device time of this rank's pull (CUDA events on the compute stream).
int dfsaLaunchPull(dfsa_state* s, const double2* remote);
extern "C" int dfsa_xk_measure_link(dfsa_state* s, int pairRank, int mode, double* ms) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s && ms && mode >= 0 && mode <= 2, "bad argument");
    DFSA_TRY(checkXArgs(s, DFSA_AMPS, 0, DFSA_BUFFER, 0, s->numAmps, pairRank));
    DFSA_REQUIRE(fusedAvailable(), "peer shards are not mapped (staged transport)");
    DfsaContext& c = dfsaCtx();
    cudaEvent_t e0, e1;
    DFSA_CUDA(cudaEventCreate(&e0));
    DFSA_CUDA(cudaEventCreate(&e1));
    const size_t bytes = s->numAmps * sizeof(double2);
    int rc = fusedGroupExchange(s, &pairRank, 1,
        [&](const double2* const* remote) -> int {
            DFSA_CUDA(cudaEventRecord(e0, c.compute));
            if (mode == 2 && (c.rank & 1)) { /* one-way: odd ranks only serve */ }
            else if (mode == 0 || mode == 2) DFSA_TRY(dfsaLaunchPull(s, remote[0]));
            else DFSA_CUDA(cudaMemcpyAsync(s->arr[DFSA_BUFFER], remote[0], bytes, cudaMemcpyDeviceToDevice, c.compute));
            DFSA_CUDA(cudaEventRecord(e1, c.compute));
            return (int)DFSA_OK;
        },
        [](dfsa_state*) { return (int)DFSA_OK; });                      // the pulled copy is discarded: no array swap
    if (rc == DFSA_OK) {
        float f = 0.f;
        if (cudaEventSynchronize(e1) != cudaSuccess || cudaEventElapsedTime(&f, e0, e1) != cudaSuccess) { dfsaSetError("link timing failed"); rc = DFSA_ERR_CUDA; }
        *ms = f;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
}

// manyCtrlOneTargGate, prefix target + suffix controls (distributed_statevector.hpp:43-78)
extern "C" int dfsa_xk_ctrlPrefixTarg(dfsa_state* s, const uint32_t* suffixCtrls, unsigned numCtrls, int pairRank, const double f0[2], const double f1[2]) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s && suffixCtrls && f0 && f1 && numCtrls >= 1 && numCtrls <= s->logNumAmps, "bad argument");
    const uint64_t m = s->numAmps >> numCtrls, allOnes = (1ULL << numCtrls) - 1ULL;
    DFSA_TRY(checkXArgs(s, DFSA_BUFFER, 0, DFSA_BUFFER, m, m, pairRank));
    if (fusedAvailable()) {
        BitSpec spec;
        uint64_t ones;
        for (unsigned q = 1; q < numCtrls; q++) DFSA_REQUIRE(suffixCtrls[q] > suffixCtrls[q - 1], "controls must be strictly increasing");
        DFSA_TRY(sortedSpec(suffixCtrls, numCtrls, s->logNumAmps, &spec, &ones));
        const double2 c0 = make_double2(f0[0], f0[1]), c1 = make_double2(f1[0], f1[1]);
        return fusedGroupExchange(s, &pairRank, 1,
            [&](const double2* const* remote) { return dfsaLaunchFusedCombineSub(s, remote[0], spec, ones, c0, c1, dfsaCtx().rank > pairRank); },
            [&](dfsa_state* st) { return dfsa_k_unpack(st, suffixCtrls, numCtrls, allOnes, 0); });      // both ranks are done reading: results go home
    }
    DFSA_TRY(dfsa_k_pack(s, suffixCtrls, numCtrls, allOnes, 0));
    DFSA_TRY(transfer(s, DFSA_BUFFER, 0, DFSA_BUFFER, m, m, pairRank, true, true));
    return dfsaCombineSub(s, suffixCtrls, numCtrls, allOnes, m, f0, f1, dfsaCtx().rank > pairRank);
}

// twoQubitDepolarising, qb1 suffix / qb2 prefix (distributed_densitymatrix.hpp:146-183)
extern "C" int dfsa_xk_depol2Pair(dfsa_state* s, unsigned qb1, unsigned qb2, unsigned bit, double prob, int corrected, int pairRank) {
    DFSA_TRY(dfsaEnsureDevice());
    const uint64_t eighth = s ? s->numAmps >> 3 : 0;
    DFSA_TRY(checkXArgs(s, DFSA_BUFFER, 0, DFSA_BUFFER, eighth, eighth, pairRank));
    DFSA_REQUIRE(s->isDensity && qb1 < qb2 && qb2 < s->numQubits && qb2 >= s->numQubits - s->logNumNodes && qb1 < s->numQubits - s->logNumNodes,
                 "needs a density matrix, qb1 suffix and qb2 prefix");
    const unsigned N = s->numQubits;
    if (fusedAvailable())
        return fusedExchange(s, pairRank, [&](const double2* remote) { return dfsaLaunchFusedDepol2Pair(s, remote, qb1, qb2, qb1 + N, bit, prob, corrected != 0); });
    const int flag = corrected ? DFSA_DEPOL2_CORRECTED : 0;
    DFSA_TRY(dfsa_k_depol2Pair(s, qb1, qb2, qb1 + N, bit, prob, 0 | flag));
    DFSA_TRY(transfer(s, DFSA_BUFFER, 0, DFSA_BUFFER, eighth, eighth, pairRank, true, true));
    return dfsa_k_depol2Pair(s, qb1, qb2, qb1 + N, bit, prob, 1 | flag);
}

// twoQubitDepolarising, both qubits prefix (distributed_densitymatrix.hpp:187-237): pairRank0 / pairRank1 differ from this rank
// in the bra bit of qb1 / qb2
extern "C" int dfsa_xk_depol2Quad(dfsa_state* s, unsigned qb1, unsigned qb2, unsigned bit0, unsigned bit1, double prob, int corrected, int pairRank0, int pairRank1) {
    DFSA_TRY(dfsaEnsureDevice());
    DfsaContext& c = dfsaCtx();
    const uint64_t quarter = s ? s->numAmps >> 2 : 0;
    DFSA_TRY(checkXArgs(s, DFSA_BUFFER, 0, DFSA_BUFFER, quarter, quarter, pairRank0));
    DFSA_TRY(checkXArgs(s, DFSA_BUFFER, 0, DFSA_BUFFER, quarter, quarter, pairRank1));
    DFSA_REQUIRE(s->isDensity && qb1 < qb2 && qb2 < s->numQubits && qb1 >= s->numQubits - s->logNumNodes, "needs a density matrix and two prefix qubits");
    if (fusedAvailable()) {
        const int peers[3] = {pairRank0, pairRank1, pairRank0 ^ pairRank1 ^ c.rank};
        return fusedGroupExchange(s, peers, 3, [&](const double2* const* remote) { return dfsaLaunchFusedDepol2Quad(s, remote, qb1, qb2, bit0, bit1, prob, corrected != 0); });
    }
    const int flag = corrected ? DFSA_DEPOL2_CORRECTED : 0;
    DFSA_TRY(dfsa_k_depol2Quad(s, qb1, qb2, bit0, bit1, prob, 0 | flag));
    DFSA_TRY(transfer(s, DFSA_BUFFER, 0, DFSA_BUFFER, quarter, quarter, pairRank0, true, true));
    DFSA_TRY(dfsa_k_depol2Quad(s, qb1, qb2, bit0, bit1, prob, 1 | flag));
    DFSA_TRY(transfer(s, DFSA_BUFFER, 0, DFSA_BUFFER, quarter, quarter, pairRank1, true, true));
    return dfsa_k_depol2Quad(s, qb1, qb2, bit0, bit1, prob, 2 | flag);
}

extern "C" int dfsa_xk_exchangePauliCombine(dfsa_state* s, int pairRank, uint64_t maskXY, uint64_t maskYZ, unsigned numY,
                                            const double f[2], const double g[2], int exact) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(f && g, "null factor");
    DFSA_TRY(checkXArgs(s, DFSA_AMPS, 0, DFSA_BUFFER, 0, s ? s->numAmps : 0, pairRank));
    const double2 ff = make_double2(f[0], f[1]), gg = make_double2(g[0], g[1]), pw = dfsaPowIHost(numY);
    const double2 h = make_double2(gg.x * pw.x - gg.y * pw.y, gg.x * pw.y + gg.y * pw.x);
    if (fusedAvailable())
        return fusedExchange(s, pairRank, [&](const double2* remote) {
            return dfsaLaunchFusedPauliCombine(s, remote, pairRank, maskXY, maskYZ, numY, ff, h, exact != 0);
        });
    int chunks = (dfsaCtx().transport == Transport::Nccl) ? chunkCountFor(s->numAmps) : 1;
    // the combine of chunk k reads buffer[j ^ maskXY]: that stays inside chunk k only while maskXY < chunk size
    while (chunks > 1 && maskXY >= s->numAmps / chunks) chunks >>= 1;
    if (chunks == 1) {
        DFSA_TRY(transfer(s, DFSA_AMPS, 0, DFSA_BUFFER, 0, s->numAmps, pairRank, true, true));
        return dfsaLaunchPauliCombineRange(s, 0, s->numAmps, pairRank, maskXY, maskYZ, numY, ff, h, exact != 0);
    }
    return pipelinedExchange(s, pairRank, chunks, [&](uint64_t first, uint64_t num) {
        return dfsaLaunchPauliCombineRange(s, first, num, pairRank, maskXY, maskYZ, numY, ff, h, exact != 0);
    });
}

// Host values reduced over the ranks (n <= 4; sum in rank order -- deterministic, identical on every rank -- or max with
// NaN propagation). Through the shared page; on the id-only NCCL bootstrap through ncclAllReduce.
int dfsaAllreduceDoubles(double* v, int n, bool isMax) {
    DfsaContext& c = dfsaCtx();
    DFSA_REQUIRE(v && n >= 1 && n <= 4, "bad argument");
    if (c.size == 1) return DFSA_OK;
    if (g_comm.shm) {
        for (int i = 0; i < n; i++) g_comm.shm->reduce[c.rank][i] = v[i];
        DFSA_TRY(shmBarrier());
        double acc[4];
        for (int i = 0; i < n; i++) {
            acc[i] = g_comm.shm->reduce[0][i];
            for (int r = 1; r < c.size; r++) {
                const double x = g_comm.shm->reduce[r][i];
                if (!isMax) acc[i] += x;
                else if (x != x || acc[i] != acc[i]) acc[i] = x != x ? x : acc[i];
                else if (x > acc[i]) acc[i] = x;
            }
        }
        DFSA_TRY(shmBarrier());
        for (int i = 0; i < n; i++) v[i] = acc[i];
        return DFSA_OK;
    }
    double2* scratch;
    DFSA_TRY(dfsaScratch(64, &scratch));
    DFSA_CUDA(cudaStreamSynchronize(c.compute));
    DFSA_CUDA(cudaMemcpyAsync(scratch, v, n * sizeof(double), cudaMemcpyHostToDevice, c.comm));
    DFSA_NCCL(ncclAllReduce(scratch, scratch, n, ncclDouble, isMax ? ncclMax : ncclSum, g_comm.nccl, c.comm));
    DFSA_CUDA(cudaMemcpyAsync(v, scratch, n * sizeof(double), cudaMemcpyDeviceToHost, c.comm));
    DFSA_CUDA(cudaStreamSynchronize(c.comm));
    return DFSA_OK;
}

extern "C" int dfsa_x_allreduce_amp(double reim[2]) {
    DFSA_REQUIRE(reim, "null argument");
    return dfsaAllreduceDoubles(reim, 2, false);
}

// ---- expecPauliString: where the reduction kernel publishes its result, and how the host collects it
namespace {
bool      g_expecLocalOnly = false;
unsigned* g_expecTicket = nullptr;        // device: block ticket of the last-block reduction
double*   g_expecPinnedDev = nullptr;     // device view of the context's pinned page
uint64_t  g_expecLocalSeq = 0;
bool      g_expecLastGlobal = false;
}

int dfsaExpecLocalOnly(bool on) { g_expecLocalOnly = on; return DFSA_OK; }

int dfsaExpecTarget(double** value, unsigned long long** flag, unsigned long long* seq, unsigned** ticket, int* global) {
    DfsaContext& c = dfsaCtx();
    if (!g_expecTicket) {
        DFSA_CUDA(cudaMalloc((void**)&g_expecTicket, 256));
        DFSA_CUDA(cudaMemsetAsync(g_expecTicket, 0, 256, c.compute));
        DFSA_CUDA(cudaHostGetDevicePointer((void**)&g_expecPinnedDev, c.hostPinned, 0));
    }
    *ticket = g_expecTicket;
    if (c.size > 1 && g_comm.signals && !g_expecLocalOnly) {
        const uint64_t q = ++g_comm.expecSeq;
        char* base = (char*)g_comm.shmDev;
        *value = (double*)(base + offsetof(Shm, expecVal) + sizeof(double) * 2 * ((q & 1) * MAXP + c.rank));
        *flag = (unsigned long long*)(base + offsetof(Shm, expecFlag) + sizeof(uint64_t) * ((q & 1) * MAXP + c.rank));
        *seq = q;
        *global = 1;
        g_expecLastGlobal = true;
        return DFSA_OK;
    }
    const uint64_t q = ++g_expecLocalSeq;
    *value = g_expecPinnedDev + 8;                              // doubles 8, 9 of the pinned page; flag in slot 10
    *flag = (unsigned long long*)(g_expecPinnedDev + 10);
    *seq = q;
    *global = 0;
    g_expecLastGlobal = false;
    return DFSA_OK;
}

static int spinForFlag(volatile uint64_t* flag, uint64_t seq, const char* what) {
    DfsaContext& c = dfsaCtx();
    const double t0 = nowSeconds();
    for (uint64_t spins = 0; *flag < seq; spins++) {
        if ((spins & 0xFFF) == 0xFFF) {
            cudaError_t e = cudaStreamQuery(c.compute);
            if (e != cudaSuccess && e != cudaErrorNotReady) { dfsaSetError("expecPauliString kernel failed: %s", cudaGetErrorString(e)); return DFSA_ERR_CUDA; }
            if (nowSeconds() - t0 > commTimeoutSeconds()) { dfsaSetError("timed out waiting for %s", what); return DFSA_ERR_COMM; }
        }
    }
    __sync_synchronize();
    return DFSA_OK;
}

int dfsaExpecCollect(unsigned long long seq, double out[2]) {
    DfsaContext& c = dfsaCtx();
    if (!g_expecLastGlobal) {
        volatile uint64_t* flag = (volatile uint64_t*)(c.hostPinned + 10);
        DFSA_TRY(spinForFlag(flag, seq, "the reduction kernel"));
        out[0] = ((volatile double*)c.hostPinned)[8];
        out[1] = ((volatile double*)c.hostPinned)[9];
        return DFSA_OK;
    }
    // every rank reads every rank's slot and sums in rank order: deterministic and identical everywhere, no barrier. A slot is
    // rewritten two calls later, which needs every rank to have posted the call in between, i.e. to have finished reading this one.
    Shm* m = g_comm.shm;
    double re = 0.0, im = 0.0;
    for (int r = 0; r < c.size; r++) {
        DFSA_TRY(spinForFlag(&m->expecFlag[seq & 1][r], seq, "a rank's expectation value"));
        re += m->expecVal[seq & 1][r][0];
        im += m->expecVal[seq & 1][r][1];
    }
    out[0] = re; out[1] = im;
    return DFSA_OK;
}

// getAllVecAmps (tests/test_utilities.hpp:419-435): every rank ends up with the whole state in host memory
extern "C" int dfsa_state_download_all(dfsa_state* s, double* hostAll) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s && hostAll, "null argument");
    DfsaContext& c = dfsaCtx();
    DFSA_TRY(dfsa_comm_barrier());
    if (c.size == 1) return dfsa_state_download(s, DFSA_AMPS, 0, s->numAmps, hostAll);
    size_t shardBytes = s->numAmps * sizeof(double2);
    if (c.transport == Transport::Nccl && !(g_comm.shm && g_comm.peersOk)) {
        // no peer mappings: shard by shard through the exchange buffer (nothing else is in flight after the barrier above), so
        // the gather needs no device memory beyond what the state already owns
        for (int r = 0; r < c.size; r++) {
            DFSA_NCCL(ncclBroadcast(s->arr[DFSA_AMPS], s->arr[DFSA_BUFFER], 2 * s->numAmps, ncclDouble, r, g_comm.nccl, c.comm));
            DFSA_CUDA(cudaMemcpyAsync((char*)hostAll + shardBytes * r, r == c.rank ? s->arr[DFSA_AMPS] : s->arr[DFSA_BUFFER], shardBytes, cudaMemcpyDeviceToHost, c.comm));
            DFSA_CUDA(cudaStreamSynchronize(c.comm));
        }
    } else {
        for (int r = 0; r < c.size; r++) {
            double2* src = s->arr[DFSA_AMPS];
            if (r != c.rank) DFSA_TRY(peerArray(s, r, DFSA_AMPS, &src));
            DFSA_CUDA(cudaMemcpyAsync((char*)hostAll + shardBytes * r, src, shardBytes, cudaMemcpyDeviceToHost, c.comm));
        }
        DFSA_CUDA(cudaStreamSynchronize(c.comm));
    }
    return dfsaHostBarrier();
}
