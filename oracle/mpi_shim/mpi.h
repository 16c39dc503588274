/*
 * mpi.h -- single-machine stand-in for the 20 MPI symbols the reference uses
 * (reference: src/communication.hpp:8-177, tests/test_utilities.hpp:429).
 *
 * TEST INFRASTRUCTURE ONLY. There is no MPI in this image nor on the GPU box, so the
 * reference's "mpirun -np P" is emulated: MPI_Init() reads $SHIM_NP, maps one shared
 * anonymous region and fork()s P-1 children (legal because comm_init() is the first
 * statement of every reference main()). Point-to-point traffic goes through one
 * single-producer/single-consumer byte ring per ordered (src,dst) pair; Isend/Irecv only
 * queue work, and a progress pump runs inside Waitall/Barrier/Allreduce/Allgather
 * (Barrier must pump: the reference's damping path leaves an untracked Isend in flight
 * and relies on the following comm_synch(), src/distributed_densitymatrix.hpp:292,317).
 *
 * Nothing here is part of the product; it exists so that oracle/_ref can be built from
 * the unmodified reference sources with plain g++.
 */
#ifndef DFSA_MPI_SHIM_H
#define DFSA_MPI_SHIM_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <deque>
#include <vector>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
#include <sched.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;   /* value = element size in bytes */
typedef int MPI_Op;
typedef int MPI_Request;    /* index into the per-process op table */
typedef int MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_DOUBLE_COMPLEX 16
#define MPI_SUM 0
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_SUCCESS 0

namespace mpishim {

constexpr int    MAXP      = 32;
constexpr size_t RING_SIZE = size_t(1) << 20;      /* payload bytes per ring */
constexpr size_t GATHER_BYTES = size_t(1) << 26;   /* Allgather scratch (tests gather <= 2^10 amps) */

struct Ring {
    volatile uint64_t head;   /* bytes ever written  (producer-owned) */
    char pad0[56];
    volatile uint64_t tail;   /* bytes ever consumed (consumer-owned) */
    char pad1[56];
    char data[RING_SIZE];
};

struct Shared {
    volatile int barrierCount;
    volatile int barrierSense;
    double reduceSlots[MAXP][2];
    Ring rings[1];            /* P*P rings follow, then the gather area */
};

struct PendingOp {
    char*  ptr;
    size_t remaining;
    int    peer;
    bool   isSend;
    bool   done;
};

struct State {
    bool initialised = false;
    int  rank = 0, size = 1;
    Shared* shm = nullptr;
    char* gather = nullptr;
    int   localSense = 0;
    std::vector<pid_t> children;
    std::vector<PendingOp> ops;               /* indexed by MPI_Request */
    std::deque<int> sendQ[MAXP], recvQ[MAXP]; /* FIFO per peer: MPI non-overtaking order */
};

inline State g;

inline Ring& ring(int src, int dst) { return g.shm->rings[src * g.size + dst]; }

/* move as many bytes as currently possible; returns true if anything moved */
inline bool pumpOnce() {
    bool moved = false;
    for (int peer = 0; peer < g.size; peer++) {
        while (!g.sendQ[peer].empty()) {
            PendingOp& op = g.ops[g.sendQ[peer].front()];
            Ring& r = ring(g.rank, peer);
            uint64_t head = r.head, tail = r.tail;
            size_t space = RING_SIZE - (size_t)(head - tail);
            size_t n = op.remaining < space ? op.remaining : space;
            if (n > 0) {
                size_t off = head % RING_SIZE, first = RING_SIZE - off;
                if (first > n) first = n;
                memcpy(r.data + off, op.ptr, first);
                memcpy(r.data, op.ptr + first, n - first);
                __sync_synchronize();
                r.head = head + n;
                op.ptr += n; op.remaining -= n; moved = true;
            }
            if (op.remaining == 0) { op.done = true; g.sendQ[peer].pop_front(); }
            else break;
        }
        while (!g.recvQ[peer].empty()) {
            PendingOp& op = g.ops[g.recvQ[peer].front()];
            Ring& r = ring(peer, g.rank);
            uint64_t head = r.head, tail = r.tail;
            __sync_synchronize();
            size_t avail = (size_t)(head - tail);
            size_t n = op.remaining < avail ? op.remaining : avail;
            if (n > 0) {
                size_t off = tail % RING_SIZE, first = RING_SIZE - off;
                if (first > n) first = n;
                memcpy(op.ptr, r.data + off, first);
                memcpy(op.ptr + first, r.data, n - first);
                __sync_synchronize();
                r.tail = tail + n;
                op.ptr += n; op.remaining -= n; moved = true;
            }
            if (op.remaining == 0) { op.done = true; g.recvQ[peer].pop_front(); }
            else break;
        }
    }
    return moved;
}

inline bool sendsOutstanding() {
    for (int p = 0; p < g.size; p++) if (!g.sendQ[p].empty()) return true;
    return false;
}

inline void barrier() {
    /* flush our own queued sends first (someone may be blocked receiving them) */
    while (sendsOutstanding()) if (!pumpOnce()) sched_yield();
    if (g.size == 1) return;
    g.localSense = !g.localSense;
    if (__sync_add_and_fetch(&g.shm->barrierCount, 1) == g.size) {
        g.shm->barrierCount = 0;
        __sync_synchronize();
        g.shm->barrierSense = g.localSense;
    } else {
        while (g.shm->barrierSense != g.localSense) { pumpOnce(); sched_yield(); }
    }
    __sync_synchronize();
}

} /* namespace mpishim */

static inline int MPI_Initialized(int* flag) { *flag = mpishim::g.initialised; return MPI_SUCCESS; }

static inline int MPI_Init(int*, char***) {
    using namespace mpishim;
    const char* np = getenv("SHIM_NP");
    g.size = np ? atoi(np) : 1;
    if (g.size < 1 || g.size > MAXP) { fprintf(stderr, "mpi shim: bad SHIM_NP\n"); exit(2); }
    size_t bytes = sizeof(Shared) + sizeof(Ring) * (size_t)g.size * g.size + GATHER_BYTES;
    void* mem = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (mem == MAP_FAILED) { perror("mpi shim mmap"); exit(2); }
    g.shm = (Shared*)mem;
    g.gather = (char*)mem + sizeof(Shared) + sizeof(Ring) * (size_t)g.size * g.size;
    g.shm->barrierCount = 0; g.shm->barrierSense = 0;
    fflush(stdout); fflush(stderr);
    for (int r = 1; r < g.size; r++) {
        pid_t pid = fork();
        if (pid < 0) { perror("mpi shim fork"); exit(2); }
        if (pid == 0) {
            g.rank = r; g.children.clear();
            if (!getenv("SHIM_KEEP_STDOUT")) { if (!freopen("/dev/null", "w", stdout)) {} }
            break;
        }
        g.children.push_back(pid);
    }
    /* what mpirun --bind-to would do: give each rank its own contiguous slice of the allowed cores,
     * so that OMP_PROC_BIND=true (reference compile.sh:13) does not pile every rank onto core 0 */
    if (g.size > 1 && !getenv("SHIM_NO_BIND")) {
        cpu_set_t allowed, mine;
        if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
            std::vector<int> cpus;
            for (int c = 0; c < CPU_SETSIZE; c++) if (CPU_ISSET(c, &allowed)) cpus.push_back(c);
            int per = (int)cpus.size() / g.size;
            if (per >= 1) {
                CPU_ZERO(&mine);
                for (int c = g.rank * per; c < (g.rank + 1) * per; c++) CPU_SET(cpus[c], &mine);
                sched_setaffinity(0, sizeof(mine), &mine);
            }
        }
    }
    g.initialised = true;
    return MPI_SUCCESS;
}

static inline int MPI_Finalize() {
    using namespace mpishim;
    barrier();
    fflush(stdout);
    if (g.rank != 0) _exit(0);     /* children never return into the caller's main() epilogue */
    for (pid_t pid : g.children) { int st; waitpid(pid, &st, 0); }
    return MPI_SUCCESS;
}

static inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = mpishim::g.rank; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm, int* s) { *s = mpishim::g.size; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm) { mpishim::barrier(); return MPI_SUCCESS; }

static inline int MPI_Isend(const void* buf, unsigned long long count, MPI_Datatype dt, int dest, int, MPI_Comm, MPI_Request* req) {
    using namespace mpishim;
    g.ops.push_back(PendingOp{(char*)buf, (size_t)count * dt, dest, true, false});
    *req = (int)g.ops.size() - 1;
    g.sendQ[dest].push_back(*req);
    return MPI_SUCCESS;
}

static inline int MPI_Irecv(void* buf, unsigned long long count, MPI_Datatype dt, int src, int, MPI_Comm, MPI_Request* req) {
    using namespace mpishim;
    g.ops.push_back(PendingOp{(char*)buf, (size_t)count * dt, src, false, false});
    *req = (int)g.ops.size() - 1;
    g.recvQ[src].push_back(*req);
    return MPI_SUCCESS;
}

static inline int MPI_Waitall(int n, MPI_Request* reqs, MPI_Status*) {
    using namespace mpishim;
    for (;;) {
        bool all = true;
        for (int i = 0; i < n; i++) if (!g.ops[reqs[i]].done) { all = false; break; }
        if (all) break;
        if (!pumpOnce()) sched_yield();
    }
    /* recycle the table once nothing is pending (untracked sends keep it alive) */
    bool idle = !sendsOutstanding();
    for (int p = 0; idle && p < g.size; p++) if (!g.recvQ[p].empty()) idle = false;
    if (idle) g.ops.clear();
    return MPI_SUCCESS;
}

static inline int MPI_Allreduce(const void* in, void* out, int count, MPI_Datatype dt, MPI_Op, MPI_Comm) {
    using namespace mpishim;
    if (count != 1 || dt != MPI_DOUBLE_COMPLEX) { fprintf(stderr, "mpi shim: Allreduce supports 1 complex double\n"); exit(2); }
    memcpy((void*)g.shm->reduceSlots[g.rank], in, 16);
    barrier();
    double re = 0, im = 0;
    for (int r = 0; r < g.size; r++) { re += g.shm->reduceSlots[r][0]; im += g.shm->reduceSlots[r][1]; }
    barrier();
    ((double*)out)[0] = re; ((double*)out)[1] = im;
    return MPI_SUCCESS;
}

static inline int MPI_Allgather(const void* in, unsigned long long count, MPI_Datatype dt, void* out, unsigned long long, MPI_Datatype, MPI_Comm) {
    using namespace mpishim;
    size_t bytes = (size_t)count * dt;
    if (bytes * g.size > GATHER_BYTES) { fprintf(stderr, "mpi shim: Allgather too large\n"); exit(2); }
    memcpy(g.gather + bytes * g.rank, in, bytes);
    barrier();
    memcpy(out, g.gather, bytes * g.size);
    barrier();
    return MPI_SUCCESS;
}

#endif /* DFSA_MPI_SHIM_H */
