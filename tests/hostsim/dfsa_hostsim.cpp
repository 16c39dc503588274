// dfsa_hostsim.cpp -- TEST INFRASTRUCTURE ONLY: a CPU stand-in for the C-ABI of include/dfsa_b200.h.
//
// What it is for: the drop-in host layer (distributed-full-state-algorithms_b200/host/*.hpp) takes every decision of the
// distributed algorithms on the host -- local kernel or pairwise exchange, which rank is the partner, relocation plans, the lazy
// qubit layout, which qubit a swap-in evicts, when deferred gates are launched, how a layout is restored. Those decisions can be
// wrong without any kernel being wrong, and they can be exercised without a GPU: this file implements the SEMANTICS each C-ABI
// entry documents in include/dfsa_b200.h with plain loops over shards that live in a shared-memory arena, one forked process
// per rank (DFSA_NP, like the real library's fork mode), so that tests/hostsim/fuzz.cpp can run random circuits through the
// real host headers at 1...16 ranks and compare against a dense ground truth that knows nothing about ranks. The same file
// carries the reference's own Catch2 cases (catch_on_standin), the extern "C" face of the host layer (capi_on_standin.py) and a
// toy-size dry run of bench.py's control flow (bench_dry_run.py) -- see tests/test_host_layer_fuzz.py.
//
// What it is NOT: it is not part of the product, is never built into or loaded by the package (libdfsa_b200.so /
// libdfsa_host.so), and nothing under distributed-full-state-algorithms_b200/ or bench.py refers to it. The product has no CPU
// path: without a CUDA device every entry of libdfsa_b200.so fails (tests/test_cabi_symbols.py). Only the binaries of this
// directory (tests/hostsim/_build) link it. Entries the host layer never calls (staged building blocks) and the reference's
// literal twoQubitDepolarising formulas on the prefix branches (pinned on the GPU by tests/golden) return DFSA_ERR_UNSUPPORTED.
//
// Pairwise operations synchronise PAIRWISE (a rank that fails a prefix control never enters them, reference
// distributed_statevector.hpp:92-93), collective ones with a barrier over all ranks; every wait times out with a message, so a
// host layer whose ranks disagree about who talks to whom shows up as "rank r waits for rank p", not as a hang.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <signal.h>
#include <sys/mman.h>
#include <sys/prctl.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include "dfsa_b200.h"

namespace {

using Amp = std::complex<double>;
constexpr int MAXP = 16;
constexpr uint64_t ARENA_BYTES = 8ULL << 30;        // virtual; pages are committed when touched
constexpr double WAIT_TIMEOUT_S = 30.0;

struct Ctl {
    std::atomic<int>      barCount;
    std::atomic<int>      barSense;
    std::atomic<uint64_t> pairSeq[MAXP][MAXP];
    double                reduce[MAXP][4];
    std::atomic<uint64_t> opCount[MAXP];           // how many C-ABI calls each rank has made (diagnostics)
};

struct Comm {
    bool   up = false;
    int    rank = 0, size = 1;
    Ctl*   ctl = nullptr;
    char*  arena = nullptr;
    int    localSense = 0;
    uint64_t pairCount[MAXP] = {};
    std::vector<pid_t> children;
    // deterministic arena allocator: every rank makes the same sequence of create / destroy calls (SPMD), hence the same decisions
    uint64_t bump = 0;
    std::map<uint64_t, std::vector<uint64_t>> freeBySize;
} g;

char g_err[512] = "";
void setError(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
#define SIM_REQUIRE(cond, ...) do { if (!(cond)) { setError(__VA_ARGS__); return DFSA_ERR_ARG; } } while (0)

double now() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

[[noreturn]] void die(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "hostsim[rank %d]: ", g.rank);
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    abort();
}

void barrierAll() {
    if (g.size == 1) return;
    g.localSense ^= 1;
    if (g.ctl->barCount.fetch_add(1) + 1 == g.size) {
        g.ctl->barCount.store(0);
        g.ctl->barSense.store(g.localSense);
        return;
    }
    const double t0 = now();
    while (g.ctl->barSense.load() != g.localSense) {
        if (now() - t0 > WAIT_TIMEOUT_S) die("barrier over all ranks timed out: not every rank made this collective call");
        sched_yield();
    }
}

void pairSync(int partner) {
    if (partner == g.rank) return;
    const uint64_t mine = ++g.pairCount[partner];
    g.ctl->pairSeq[g.rank][partner].store(mine);
    const double t0 = now();
    while (g.ctl->pairSeq[partner][g.rank].load() < mine) {
        if (now() - t0 > WAIT_TIMEOUT_S) die("waits for rank %d in a pairwise step that rank %d never entered", partner, partner);
        sched_yield();
    }
}

}  // namespace

struct dfsa_state {
    int      isDensity;
    unsigned numQubits, logNumAmps;
    uint64_t numAmps, base, bytesAllRanks;
    Amp* arr(int rank, int which) const { return reinterpret_cast<Amp*>(g.arena + base) + (uint64_t(rank) * 2 + which) * numAmps; }
    Amp* amps() const { return arr(g.rank, DFSA_AMPS); }
    Amp* buffer() const { return arr(g.rank, DFSA_BUFFER); }
    uint64_t rankShift() const { return uint64_t(g.rank) << logNumAmps; }
};

namespace {
inline uint64_t insertZero(uint64_t x, unsigned pos) {
    const uint64_t below = x & ((1ULL << pos) - 1ULL);
    return ((x ^ below) << 1) | below;
}
inline uint64_t insertZeros(uint64_t x, std::vector<unsigned> sortedPos) {
    for (unsigned p : sortedPos) x = insertZero(x, p);
    return x;
}
inline Amp amp2(const double v[2]) { return Amp(v[0], v[1]); }
// synthetic state: any reproducible function of (seed, GLOBAL index) will do here -- the host layer never looks at the values
inline Amp hashAmp(uint64_t seed, uint64_t globalIndex) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (globalIndex + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return Amp(double(z >> 40) / double(1 << 24) - 0.5, double((z >> 16) & 0xFFFFFF) / double(1 << 24) - 0.5);
}
inline int parity(uint64_t x) { return __builtin_parityll(x); }
unsigned long long g_remoteAmps = 0;      // amplitudes this rank has read from (or, dfsa_x_exchange, received from) other ranks' shards
std::map<std::string, unsigned long long>& callCounts() { static std::map<std::string, unsigned long long> c; return c; }
inline void touchOp(const char* entry) { if (g.ctl) g.ctl->opCount[g.rank].fetch_add(1); callCounts()[entry]++; }
#define touchOp() touchOp(__func__)
}  // namespace

extern "C" {

const char* dfsa_last_error(void) { return g_err; }
// stand-in only (declared by fuzz.cpp): how often this rank entered a C-ABI function -- lets a test assert that the paths it
// means to exercise (relocations, swap-ins, exchanges) were actually taken
// stand-in only: amplitudes this rank has pulled from other ranks so far -- what must cross NVLink in one direction for this rank,
// whatever the implementation; tests/hostsim/cost_model_check.py holds bench.py's roofline cost model against it
unsigned long long hostsim_remote_amps(void) { return g_remoteAmps; }
unsigned long long hostsim_call_count(const char* entry) { auto it = callCounts().find(entry); return it == callCounts().end() ? 0ULL : it->second; }
const char* dfsa_version(void) { return "hostsim (CPU stand-in for tests; not the product)"; }

// ---- communication environment ------------------------------------------------------------------------------------------
int dfsa_comm_init(void) {
    if (g.up) return DFSA_OK;
    const char* np = getenv("DFSA_NP");
    const int P = np ? atoi(np) : 1;
    SIM_REQUIRE(P >= 1 && P <= MAXP && (P & (P - 1)) == 0, "DFSA_NP must be a power of two <= %d", MAXP);
    void* ctl = mmap(nullptr, sizeof(Ctl), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    void* arena = mmap(nullptr, ARENA_BYTES, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (ctl == MAP_FAILED || arena == MAP_FAILED) { setError("mmap failed"); return DFSA_ERR_COMM; }
    memset(ctl, 0, sizeof(Ctl));
    g.ctl = static_cast<Ctl*>(ctl);
    g.arena = static_cast<char*>(arena);
    g.size = P;
    g.rank = 0;
    fflush(stdout);
    fflush(stderr);
    for (int r = 1; r < P; r++) {
        const pid_t pid = fork();
        if (pid < 0) { setError("fork failed"); return DFSA_ERR_COMM; }
        if (pid == 0) {
            prctl(PR_SET_PDEATHSIG, SIGKILL);
            g.rank = r;
            g.children.clear();
            if (!freopen("/dev/null", "w", stdout)) {}
            break;
        }
        g.children.push_back(pid);
    }
    g.up = true;
    return DFSA_OK;
}
int dfsa_comm_get_unique_id(void*) { setError("hostsim: no NCCL"); return DFSA_ERR_UNSUPPORTED; }
int dfsa_comm_init_with_id(int, int, const void*, int) { setError("hostsim: no NCCL"); return DFSA_ERR_UNSUPPORTED; }
int dfsa_comm_finalize(void) {
    if (!g.up) return DFSA_OK;
    barrierAll();
    int bad = 0;
    for (pid_t pid : g.children) {
        int status = 0;
        waitpid(pid, &status, 0);
        if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) bad++;
    }
    g.children.clear();
    if (bad) { fprintf(stderr, "hostsim: %d rank process(es) failed\n", bad); _exit(1); }
    return DFSA_OK;
}
int dfsa_comm_rank(void) { return g.rank; }
int dfsa_comm_size(void) { return g.size; }
int dfsa_comm_barrier(void) { touchOp(); barrierAll(); return DFSA_OK; }
int dfsa_device_sync(void) { return DFSA_OK; }
const char* dfsa_comm_transport(void) { return "hostsim"; }
int dfsa_comm_set_fused(int) { return DFSA_OK; }
int dfsa_comm_fused_active(void) { return 0; }

int dfsa_x_allreduce_amp(double reim[2]) {
    touchOp();
    if (g.size == 1) return DFSA_OK;
    g.ctl->reduce[g.rank][0] = reim[0];
    g.ctl->reduce[g.rank][1] = reim[1];
    barrierAll();
    double re = 0, im = 0;
    for (int r = 0; r < g.size; r++) { re += g.ctl->reduce[r][0]; im += g.ctl->reduce[r][1]; }
    barrierAll();
    reim[0] = re;
    reim[1] = im;
    return DFSA_OK;
}

// stand-in only (bench_dry_run.py's replacement for the torch.distributed max-reduce of bench.py's Job)
int hostsim_allreduce_max(double* value) {
    if (g.size == 1) return DFSA_OK;
    g.ctl->reduce[g.rank][2] = *value;
    barrierAll();
    double m = g.ctl->reduce[0][2];
    for (int r = 1; r < g.size; r++) m = std::max(m, g.ctl->reduce[r][2]);
    barrierAll();
    *value = m;
    return DFSA_OK;
}

// ---- state storage ------------------------------------------------------------------------------------------------------
int dfsa_state_create(int isDensity, unsigned numQubits, dfsa_state** out) {
    touchOp();
    SIM_REQUIRE(g.up || dfsa_comm_init() == DFSA_OK, "comm not initialised");
    const unsigned bits = isDensity ? 2 * numQubits : numQubits;
    unsigned k = 0;
    while ((1 << k) < g.size) k++;
    SIM_REQUIRE(out && bits >= k && bits - k < 40 && (!isDensity || numQubits >= k), "state too small for %d ranks (or too large for the simulator)", g.size);
    dfsa_state* s = new dfsa_state();
    s->isDensity = isDensity;
    s->numQubits = numQubits;
    s->logNumAmps = bits - k;
    s->numAmps = 1ULL << s->logNumAmps;
    s->bytesAllRanks = uint64_t(g.size) * 2 * s->numAmps * sizeof(Amp);
    std::vector<uint64_t>& recycled = g.freeBySize[s->bytesAllRanks];
    if (!recycled.empty()) { s->base = recycled.back(); recycled.pop_back(); }
    else {
        s->base = g.bump;
        g.bump += (s->bytesAllRanks + 4095) & ~4095ULL;
        if (g.bump > ARENA_BYTES) { delete s; setError("hostsim arena exhausted"); return DFSA_ERR_CUDA; }
    }
    // nobody may still be reading a recycled region: creation is ordered after everything the ranks did before
    barrierAll();
    memset(static_cast<void*>(s->amps()), 0, s->numAmps * sizeof(Amp));
    memset(static_cast<void*>(s->buffer()), 0, s->numAmps * sizeof(Amp));
    barrierAll();
    *out = s;
    return DFSA_OK;
}
int dfsa_state_destroy(dfsa_state* s) {
    touchOp();
    if (!s) return DFSA_OK;
    g.freeBySize[s->bytesAllRanks].push_back(s->base);
    delete s;
    return DFSA_OK;
}
double*  dfsa_state_ptr(dfsa_state* s, int which) { return reinterpret_cast<double*>(s->arr(g.rank, which)); }
uint64_t dfsa_state_num_amps_per_node(const dfsa_state* s) { return s->numAmps; }
unsigned dfsa_state_log_num_amps_per_node(const dfsa_state* s) { return s->logNumAmps; }
unsigned dfsa_state_num_qubits(const dfsa_state* s) { return s->numQubits; }
int      dfsa_state_is_density(const dfsa_state* s) { return s->isDensity; }
int dfsa_state_swap_arrays(dfsa_state*) { setError("hostsim: not needed by the host layer"); return DFSA_ERR_UNSUPPORTED; }

int dfsa_state_upload(dfsa_state* s, int which, uint64_t first, uint64_t num, const double* host) {
    touchOp();
    SIM_REQUIRE(s && host && first + num <= s->numAmps, "range");
    memcpy(static_cast<void*>(s->arr(g.rank, which) + first), host, num * sizeof(Amp));
    return DFSA_OK;
}
int dfsa_state_download(dfsa_state* s, int which, uint64_t first, uint64_t num, double* host) {
    touchOp();
    SIM_REQUIRE(s && host && first + num <= s->numAmps, "range");
    memcpy(host, static_cast<void*>(s->arr(g.rank, which) + first), num * sizeof(Amp));
    return DFSA_OK;
}
int dfsa_state_download_all(dfsa_state* s, double* hostAll) {
    touchOp();
    barrierAll();
    for (int r = 0; r < g.size; r++) memcpy(hostAll + 2 * uint64_t(r) * s->numAmps, static_cast<void*>(s->arr(r, DFSA_AMPS)), s->numAmps * sizeof(Amp));
    barrierAll();
    return DFSA_OK;
}
int dfsa_state_upload_all(dfsa_state* s, const double* hostAll) {
    touchOp();
    memcpy(static_cast<void*>(s->amps()), hostAll + 2 * uint64_t(g.rank) * s->numAmps, s->numAmps * sizeof(Amp));
    barrierAll();
    return DFSA_OK;
}
int dfsa_state_init_zero(dfsa_state* s) { touchOp(); memset(static_cast<void*>(s->amps()), 0, s->numAmps * sizeof(Amp)); return DFSA_OK; }
int dfsa_state_init_hash(dfsa_state* s, uint64_t seed) {
    touchOp();
    for (uint64_t j = 0; j < s->numAmps; j++) s->amps()[j] = hashAmp(seed, s->rankShift() | j);
    return DFSA_OK;
}
int dfsa_state_init_plus(dfsa_state* s) {
    touchOp();
    const unsigned bits = s->isDensity ? 2 * s->numQubits : s->numQubits;
    const double v = s->isDensity ? std::ldexp(1.0, -int(s->numQubits)) : std::pow(2.0, -0.5 * bits);
    for (uint64_t j = 0; j < s->numAmps; j++) s->amps()[j] = Amp(v, 0);
    return DFSA_OK;
}
int dfsa_state_norm2(dfsa_state* s, double* out) {
    touchOp();
    barrierAll();
    double n = 0;
    for (int r = 0; r < g.size; r++)
        for (uint64_t j = 0; j < s->numAmps; j++) n += std::norm(s->arr(r, DFSA_AMPS)[j]);
    barrierAll();
    *out = n;
    return DFSA_OK;
}
int dfsa_state_copy(dfsa_state* dst, const dfsa_state* src) {
    touchOp();
    SIM_REQUIRE(dst && src && dst->numAmps == src->numAmps, "shape");
    memcpy(static_cast<void*>(dst->amps()), static_cast<void*>(src->amps()), src->numAmps * sizeof(Amp));
    return DFSA_OK;
}
int dfsa_state_compare(dfsa_state* a, dfsa_state* b, double* maxAbsDiff, uint64_t* numUnequal, double* maxAbsRef) {
    touchOp();
    SIM_REQUIRE(a && b && a->numAmps == b->numAmps, "shape");
    barrierAll();
    double d = 0, mr = 0;
    uint64_t ne = 0;
    for (int r = 0; r < g.size; r++)
        for (uint64_t j = 0; j < a->numAmps; j++) {
            const Amp x = a->arr(r, DFSA_AMPS)[j], y = b->arr(r, DFSA_AMPS)[j];
            d = std::max(d, std::max(std::abs(x.real() - y.real()), std::abs(x.imag() - y.imag())));
            mr = std::max(mr, std::max(std::abs(y.real()), std::abs(y.imag())));
            ne += !(x.real() == y.real() && x.imag() == y.imag());
        }
    barrierAll();
    if (maxAbsDiff) *maxAbsDiff = d;
    if (numUnequal) *numUnequal = ne;
    if (maxAbsRef) *maxAbsRef = mr;
    return DFSA_OK;
}
int dfsa_state_compare_hash(dfsa_state* s, uint64_t seed, double* maxAbsDiff, uint64_t* numUnequal, double* maxAbsRef) {
    touchOp();
    barrierAll();
    double d = 0, mr = 0;
    uint64_t ne = 0;
    for (int r = 0; r < g.size; r++)
        for (uint64_t j = 0; j < s->numAmps; j++) {
            const Amp x = s->arr(r, DFSA_AMPS)[j], y = hashAmp(seed, (uint64_t(r) << s->logNumAmps) | j);
            d = std::max(d, std::max(std::abs(x.real() - y.real()), std::abs(x.imag() - y.imag())));
            mr = std::max(mr, std::max(std::abs(y.real()), std::abs(y.imag())));
            ne += !(x.real() == y.real() && x.imag() == y.imag());
        }
    barrierAll();
    if (maxAbsDiff) *maxAbsDiff = d;
    if (numUnequal) *numUnequal = ne;
    if (maxAbsRef) *maxAbsRef = mr;
    return DFSA_OK;
}

// ---- measurement helpers of the C-ABI, so that bench.py's control flow can be dry-run on the CPU (tests/hostsim/bench_dry_run.py):
//      "events" are host time stamps, the "launch count" counts C-ABI calls, "pinned" memory is malloc
int dfsa_event_create(void** event) { *event = new double(0); return DFSA_OK; }
int dfsa_event_record(void* event) { *static_cast<double*>(event) = now(); return DFSA_OK; }
int dfsa_event_elapsed_ms(void* start, void* stop, double* ms) { *ms = std::max(1e-6, (*static_cast<double*>(stop) - *static_cast<double*>(start)) * 1e3); return DFSA_OK; }
int dfsa_event_destroy(void* event) { delete static_cast<double*>(event); return DFSA_OK; }
uint64_t dfsa_launch_count(void) { return g.ctl ? g.ctl->opCount[g.rank].load() : 0; }
int dfsa_host_alloc_pinned(uint64_t bytes, void** out) { *out = malloc(bytes); return *out ? DFSA_OK : DFSA_ERR_CUDA; }
int dfsa_host_free_pinned(void* ptr) { free(ptr); return DFSA_OK; }
void* dfsa_stream_compute(void) { return nullptr; }
int dfsa_comm_last_exchange_ms(double* ms) { *ms = -1.0; return DFSA_OK; }
int dfsa_xk_measure_link(dfsa_state*, int, int, double*) { setError("hostsim: nothing to measure"); return DFSA_ERR_UNSUPPORTED; }

// ---- pairwise exchange --------------------------------------------------------------------------------------------------
int dfsa_x_exchange(dfsa_state* s, int sendWhich, uint64_t sendStart, int recvWhich, uint64_t recvStart, uint64_t num, int pairRank) {
    touchOp();
    SIM_REQUIRE(s && pairRank >= 0 && pairRank < g.size && sendStart + num <= s->numAmps && recvStart + num <= s->numAmps, "range");
    std::vector<Amp> out(s->arr(g.rank, sendWhich) + sendStart, s->arr(g.rank, sendWhich) + sendStart + num);
    pairSync(pairRank);                                              // the partner has read what it sends
    memcpy(static_cast<void*>(s->arr(pairRank, recvWhich) + recvStart), out.data(), num * sizeof(Amp));
    g_remoteAmps += num;                                             // (the partner sends as much as it receives)
    pairSync(pairRank);                                              // what I receive has landed
    return DFSA_OK;
}
int dfsa_x_send(dfsa_state*, int, uint64_t, int, uint64_t, uint64_t, int) { setError("hostsim: dfsa_x_send not simulated"); return DFSA_ERR_UNSUPPORTED; }
int dfsa_x_recv(dfsa_state*, int, uint64_t, uint64_t, int) { setError("hostsim: dfsa_x_recv not simulated"); return DFSA_ERR_UNSUPPORTED; }

// new[j] = f0 * amps[j] + f1 * partner_amps[j]          (distributed_statevector.hpp:26-38)
int dfsa_xk_exchangeCombine(dfsa_state* s, int pairRank, const double f0[2], const double f1[2]) {
    touchOp();
    SIM_REQUIRE(s && pairRank >= 0 && pairRank < g.size && pairRank != g.rank, "partner");
    std::vector<Amp> out(s->numAmps);
    pairSync(pairRank);
    const Amp *mine = s->amps(), *theirs = s->arr(pairRank, DFSA_AMPS);
    for (uint64_t j = 0; j < s->numAmps; j++) out[j] = amp2(f0) * mine[j] + amp2(f1) * theirs[j];
    g_remoteAmps += s->numAmps;
    pairSync(pairRank);
    memcpy(static_cast<void*>(s->amps()), out.data(), s->numAmps * sizeof(Amp));
    return DFSA_OK;
}

// the same on the sub-cube where every suffix control is 1   (distributed_statevector.hpp:43-78)
int dfsa_xk_ctrlPrefixTarg(dfsa_state* s, const uint32_t* suffixCtrls, unsigned numCtrls, int pairRank, const double f0[2], const double f1[2]) {
    touchOp();
    SIM_REQUIRE(s && pairRank >= 0 && pairRank < g.size && pairRank != g.rank, "partner");
    uint64_t mask = 0;
    for (unsigned i = 0; i < numCtrls; i++) { SIM_REQUIRE(suffixCtrls[i] < s->logNumAmps && (i == 0 || suffixCtrls[i] > suffixCtrls[i - 1]), "controls must be increasing suffix bits"); mask |= 1ULL << suffixCtrls[i]; }
    std::vector<Amp> out(s->amps(), s->amps() + s->numAmps);
    pairSync(pairRank);
    const Amp *mine = s->amps(), *theirs = s->arr(pairRank, DFSA_AMPS);
    for (uint64_t j = 0; j < s->numAmps; j++)
        if ((j & mask) == mask) { out[j] = amp2(f0) * mine[j] + amp2(f1) * theirs[j]; g_remoteAmps++; }
    pairSync(pairRank);
    memcpy(static_cast<void*>(s->amps()), out.data(), s->numAmps * sizeof(Amp));
    return DFSA_OK;
}

// swap of suffix bit qb1 with the rank bit that tells this rank from pairRank   (distributed_statevector.hpp:140-186)
int dfsa_xk_swapSuffixPrefix(dfsa_state* s, unsigned qb1, unsigned movingBit, int pairRank) {
    touchOp();
    SIM_REQUIRE(s && qb1 < s->logNumAmps && movingBit <= 1 && pairRank >= 0 && pairRank < g.size && pairRank != g.rank, "arguments");
    std::vector<Amp> out(s->numAmps);
    pairSync(pairRank);
    const Amp *mine = s->amps(), *theirs = s->arr(pairRank, DFSA_AMPS);
    for (uint64_t j = 0; j < s->numAmps; j++) out[j] = (((j >> qb1) & 1ULL) != movingBit) ? mine[j] : theirs[j ^ (1ULL << qb1)];
    g_remoteAmps += s->numAmps / 2;
    pairSync(pairRank);
    memcpy(static_cast<void*>(s->amps()), out.data(), s->numAmps * sizeof(Amp));
    return DFSA_OK;
}

// COLLECTIVE: index bits suffixQubits[i] <-> prefixQubits[i] of the distributed array trade places for every i
int dfsa_xk_relocate(dfsa_state* s, const uint32_t* suffixQubits, const uint32_t* prefixQubits, unsigned numPairs) {
    touchOp();
    SIM_REQUIRE(s && numPairs >= 1 && numPairs <= 4, "1..4 pairs");
    const unsigned L = s->logNumAmps;
    unsigned bitsTotal = L;
    while ((1ULL << (bitsTotal - L)) < uint64_t(g.size)) bitsTotal++;
    uint64_t seen = 0;
    for (unsigned i = 0; i < numPairs; i++) {
        SIM_REQUIRE(suffixQubits[i] < L && prefixQubits[i] >= L && prefixQubits[i] < bitsTotal, "pair %u is not (suffix bit, rank bit)", i);
        SIM_REQUIRE(!((seen >> suffixQubits[i]) & 1ULL) && !((seen >> prefixQubits[i]) & 1ULL), "a bit occurs twice");
        seen |= (1ULL << suffixQubits[i]) | (1ULL << prefixQubits[i]);
    }
    std::vector<Amp> out(s->numAmps);
    barrierAll();
    for (uint64_t j = 0; j < s->numAmps; j++) {
        uint64_t src = s->rankShift() | j;
        for (unsigned i = 0; i < numPairs; i++) {
            const uint64_t a = (src >> suffixQubits[i]) & 1ULL, b = (src >> prefixQubits[i]) & 1ULL;
            if (a != b) src ^= (1ULL << suffixQubits[i]) | (1ULL << prefixQubits[i]);
        }
        out[j] = s->arr(int(src >> L), DFSA_AMPS)[src & (s->numAmps - 1)];
        g_remoteAmps += int(src >> L) != g.rank;
    }
    barrierAll();
    memcpy(static_cast<void*>(s->amps()), out.data(), s->numAmps * sizeof(Amp));
    return DFSA_OK;
}
int dfsa_plan_relocate(int, unsigned, const uint32_t*, unsigned, int[16], unsigned*) { setError("hostsim: host-only planners live in libdfsa_b200"); return DFSA_ERR_UNSUPPORTED; }

// new[j0] = f * amps[j0] + g * b * partner_amps[j0 ^ maskXY], b from the GLOBAL index of the amplitude read   (:227-241)
int dfsa_xk_exchangePauliCombine(dfsa_state* s, int pairRank, uint64_t maskXY, uint64_t maskYZ, unsigned numY, const double f[2], const double gg[2], int) {
    touchOp();
    SIM_REQUIRE(s && pairRank >= 0 && pairRank < g.size && pairRank != g.rank && maskXY < s->numAmps, "arguments");
    static const Amp powI[4] = {Amp(1, 0), Amp(0, 1), Amp(-1, 0), Amp(0, -1)};
    std::vector<Amp> out(s->numAmps);
    pairSync(pairRank);
    const Amp *mine = s->amps(), *theirs = s->arr(pairRank, DFSA_AMPS);
    for (uint64_t j0 = 0; j0 < s->numAmps; j0++) {
        const uint64_t j1 = j0 ^ maskXY, global1 = (uint64_t(pairRank) << s->logNumAmps) | j1;
        const Amp b = powI[numY & 3u] * (parity(global1 & maskYZ) ? -1.0 : 1.0);
        out[j0] = amp2(f) * mine[j0] + amp2(gg) * b * theirs[j1];
        g_remoteAmps++;
    }
    pairSync(pairRank);
    memcpy(static_cast<void*>(s->amps()), out.data(), s->numAmps * sizeof(Amp));
    return DFSA_OK;
}


// ---- rank-local kernels -------------------------------------------------------------------------------------------------
static int applyCtrlOneTarg(dfsa_state* s, uint64_t localCtrlMask, unsigned target, const double gate[8]) {
    const Amp m00 = amp2(gate), m01 = amp2(gate + 2), m10 = amp2(gate + 4), m11 = amp2(gate + 6);
    Amp* a = s->amps();
    const uint64_t t = 1ULL << target;
    for (uint64_t j = 0; j < s->numAmps; j++) {
        if ((j & t) || (j & localCtrlMask) != localCtrlMask) continue;
        const Amp a0 = a[j], a1 = a[j | t];
        a[j] = m00 * a0 + m01 * a1;
        a[j | t] = m10 * a0 + m11 * a1;
    }
    return DFSA_OK;
}
int dfsa_k_ctrlOneTarg(dfsa_state* s, const uint32_t* ctrls, unsigned numCtrls, unsigned target, const double gate[8]) {
    touchOp();
    SIM_REQUIRE(s && target < s->logNumAmps, "target must be a suffix bit");
    uint64_t mask = 0;
    for (unsigned i = 0; i < numCtrls; i++) { SIM_REQUIRE(ctrls[i] < s->logNumAmps && ctrls[i] != target, "controls must be suffix bits other than the target"); mask |= 1ULL << ctrls[i]; }
    return applyCtrlOneTarg(s, mask, target, gate);
}
int dfsa_k_gateSequence(dfsa_state* s, const dfsa_gate1* gates, unsigned numGates) {
    touchOp();
    const uint64_t localMask = s->numAmps - 1;
    for (unsigned i = 0; i < numGates; i++) {
        SIM_REQUIRE(gates[i].target < s->logNumAmps, "targets must be suffix bits");
        SIM_REQUIRE(!((gates[i].ctrlMask >> gates[i].target) & 1ULL), "a gate cannot be controlled on its own target");
        const uint64_t pre = gates[i].ctrlMask & ~localMask;
        if ((s->rankShift() & pre) != pre) continue;                 // a control on a rank bit this rank fails
        applyCtrlOneTarg(s, gates[i].ctrlMask & localMask, gates[i].target, gates[i].matrix);
    }
    return DFSA_OK;
}
int dfsa_plan_gateSequence(const dfsa_gate1*, unsigned, unsigned, uint32_t*, uint32_t*, uint32_t*, uint32_t*, uint32_t*, unsigned*) { setError("hostsim: host-only planners live in libdfsa_b200"); return DFSA_ERR_UNSUPPORTED; }

int dfsa_k_swap(dfsa_state* s, unsigned qb1, unsigned qb2) {
    touchOp();
    SIM_REQUIRE(s && qb1 < s->logNumAmps && qb2 < s->logNumAmps && qb1 != qb2, "two distinct suffix bits");
    Amp* a = s->amps();
    for (uint64_t j = 0; j < s->numAmps; j++)
        if (((j >> qb1) & 1ULL) && !((j >> qb2) & 1ULL)) std::swap(a[j], a[j ^ (1ULL << qb1) ^ (1ULL << qb2)]);
    return DFSA_OK;
}

static int applyManyTarg(dfsa_state* s, const uint32_t* targets, unsigned t, const Amp* gate) {
    const uint64_t D = 1ULL << t;
    std::vector<unsigned> sorted(targets, targets + t);
    std::sort(sorted.begin(), sorted.end());
    for (unsigned i = 0; i < t; i++) SIM_REQUIRE(sorted[i] < s->logNumAmps && (i == 0 || sorted[i] != sorted[i - 1]), "targets must be distinct suffix bits");
    std::vector<uint64_t> offset(D, 0);
    for (uint64_t r = 0; r < D; r++)
        for (unsigned i = 0; i < t; i++) if ((r >> i) & 1ULL) offset[r] |= 1ULL << targets[i];
    std::vector<Amp> v(D);
    Amp* a = s->amps();
    for (uint64_t m = 0; m < (s->numAmps >> t); m++) {
        const uint64_t base = insertZeros(m, sorted);
        for (uint64_t c = 0; c < D; c++) v[c] = a[base | offset[c]];
        for (uint64_t r = 0; r < D; r++) {
            Amp acc(0, 0);
            for (uint64_t c = 0; c < D; c++) acc += gate[r * D + c] * v[c];
            a[base | offset[r]] = acc;
        }
    }
    return DFSA_OK;
}
int dfsa_k_manyTarg(dfsa_state* s, const uint32_t* targets, unsigned numTargets, const double* gate) {
    touchOp();
    SIM_REQUIRE(s && numTargets >= 1 && numTargets <= s->logNumAmps && numTargets <= 12, "target count");
    return applyManyTarg(s, targets, numTargets, reinterpret_cast<const Amp*>(gate));
}
// sum_K conj(K) (x) K on {targets, targets + N}: row = i*d + k, col = j*d + l   (misc.hpp:58-81)
int dfsa_k_krausMap(dfsa_state* s, const uint32_t* targets2t, unsigned numTargets2t, const double* krausOps, unsigned numOps) {
    touchOp();
    SIM_REQUIRE(s && numTargets2t >= 2 && !(numTargets2t & 1u) && numTargets2t <= 12 && numOps >= 1, "arguments");
    const uint64_t d = 1ULL << (numTargets2t / 2), D = d * d;
    std::vector<Amp> super(D * D, Amp(0, 0));
    const Amp* K = reinterpret_cast<const Amp*>(krausOps);
    for (unsigned o = 0; o < numOps; o++, K += d * d)
        for (uint64_t i = 0; i < d; i++)
            for (uint64_t j = 0; j < d; j++)
                for (uint64_t k = 0; k < d; k++)
                    for (uint64_t l = 0; l < d; l++) super[(i * d + k) * D + (j * d + l)] += std::conj(K[i * d + j]) * K[k * d + l];
    return applyManyTarg(s, targets2t, numTargets2t, super.data());
}
int dfsa_plan_manyTargLayout(const uint32_t*, unsigned, unsigned, uint32_t[9], uint32_t[9], uint32_t[9], uint32_t[6], uint32_t[6]) { setError("hostsim: host-only planners live in libdfsa_b200"); return DFSA_ERR_UNSUPPORTED; }

int dfsa_k_pauli(dfsa_state* s, uint64_t maskXY, uint64_t maskYZ, unsigned numY, const double f[2], const double gg[2], int) {
    touchOp();
    SIM_REQUIRE(s && maskXY < s->numAmps, "X/Y targets must be suffix bits");
    static const Amp powI[4] = {Amp(1, 0), Amp(0, 1), Amp(-1, 0), Amp(0, -1)};
    std::vector<Amp> out(s->numAmps);
    const Amp* a = s->amps();
    for (uint64_t j0 = 0; j0 < s->numAmps; j0++) {
        const uint64_t j1 = j0 ^ maskXY;
        const Amp b = powI[numY & 3u] * (parity((s->rankShift() | j1) & maskYZ) ? -1.0 : 1.0);
        out[j0] = amp2(f) * a[j0] + amp2(gg) * b * a[j1];
    }
    memcpy(static_cast<void*>(s->amps()), out.data(), s->numAmps * sizeof(Amp));
    return DFSA_OK;
}
int dfsa_k_phase(dfsa_state* s, uint64_t targMask, double theta) {
    touchOp();
    const Amp even(std::cos(theta), std::sin(theta)), odd(std::cos(theta), -std::sin(theta));
    for (uint64_t j = 0; j < s->numAmps; j++) s->amps()[j] *= parity((s->rankShift() | j) & targMask) ? odd : even;
    return DFSA_OK;
}
int dfsa_k_scaleAll(dfsa_state* s, const double factor[2]) {
    touchOp();
    for (uint64_t j = 0; j < s->numAmps; j++) s->amps()[j] *= amp2(factor);
    return DFSA_OK;
}
int dfsa_k_copyFromBuffer(dfsa_state* s, uint64_t dstStart, uint64_t srcStart, uint64_t num) {
    touchOp();
    SIM_REQUIRE(s && dstStart + num <= s->numAmps && srcStart + num <= s->numAmps, "range");
    memcpy(static_cast<void*>(s->amps() + dstStart), static_cast<void*>(s->buffer() + srcStart), num * sizeof(Amp));
    return DFSA_OK;
}
int dfsa_k_combine(dfsa_state*, const double[2], const double[2]) { setError("hostsim: staged building block, not used by the host layer"); return DFSA_ERR_UNSUPPORTED; }
int dfsa_k_pack(dfsa_state*, const uint32_t*, unsigned, uint64_t, uint64_t) { setError("hostsim: staged building block, not used by the host layer"); return DFSA_ERR_UNSUPPORTED; }
int dfsa_k_unpack(dfsa_state*, const uint32_t*, unsigned, uint64_t, uint64_t) { setError("hostsim: staged building block, not used by the host layer"); return DFSA_ERR_UNSUPPORTED; }
int dfsa_k_combineSub(dfsa_state*, const uint32_t*, unsigned, uint64_t, uint64_t, const double[2], const double[2]) { setError("hostsim: staged building block, not used by the host layer"); return DFSA_ERR_UNSUPPORTED; }
int dfsa_k_pauliCombine(dfsa_state*, int, uint64_t, uint64_t, unsigned, const double[2], const double[2], int) { setError("hostsim: staged building block, not used by the host layer"); return DFSA_ERR_UNSUPPORTED; }

// ---- density-matrix kernels: the two dephasing channels (they never communicate) and the local partial trace ------------------
int dfsa_k_oneQubitDephasing(dfsa_state* s, unsigned qb, double prob) {
    touchOp();
    SIM_REQUIRE(s && s->isDensity && qb < s->numQubits, "arguments");
    for (uint64_t j = 0; j < s->numAmps; j++) {
        const uint64_t i = s->rankShift() | j;
        if (((i >> qb) ^ (i >> (qb + s->numQubits))) & 1ULL) s->amps()[j] *= 1 - 2 * prob;
    }
    return DFSA_OK;
}
int dfsa_k_twoQubitDephasing(dfsa_state* s, unsigned qb1, unsigned qb2, double prob) {
    touchOp();
    SIM_REQUIRE(s && s->isDensity && qb1 < s->numQubits && qb2 < s->numQubits && qb1 != qb2, "arguments");
    const unsigned N = s->numQubits;
    for (uint64_t j = 0; j < s->numAmps; j++) {
        const uint64_t i = s->rankShift() | j;
        if ((((i >> qb1) ^ (i >> (qb1 + N))) | ((i >> qb2) ^ (i >> (qb2 + N)))) & 1ULL) s->amps()[j] *= 1 - 4 * prob / 3;
    }
    return DFSA_OK;
}
// out[l] = sum_k in[l with k on `targets` and k on `pairTargets`], all traced bits rank-local   (local_densitymatrix.hpp:134-164)
int dfsa_k_partialTrace(dfsa_state* in, dfsa_state* out, const uint32_t* targets, const uint32_t* pairTargets, unsigned numTargets) {
    touchOp();
    SIM_REQUIRE(in && out && in->isDensity && out->isDensity && numTargets >= 1 && out->numQubits + numTargets == in->numQubits, "shapes");
    SIM_REQUIRE(out->numAmps << (2 * numTargets) == in->numAmps, "every traced bit must be a suffix bit of the input");
    std::vector<unsigned> all(targets, targets + numTargets);
    all.insert(all.end(), pairTargets, pairTargets + numTargets);
    std::sort(all.begin(), all.end());
    for (unsigned i = 0; i < all.size(); i++) SIM_REQUIRE(all[i] < in->logNumAmps && (i == 0 || all[i] != all[i - 1]), "traced bits must be distinct suffix bits");
    for (uint64_t l = 0; l < out->numAmps; l++) {
        const uint64_t base = insertZeros(l, all);
        Amp acc(0, 0);
        for (uint64_t k = 0; k < (1ULL << numTargets); k++) {
            uint64_t idx = base;
            for (unsigned i = 0; i < numTargets; i++) if ((k >> i) & 1ULL) idx |= (1ULL << targets[i]) | (1ULL << pairTargets[i]);
            acc += in->amps()[idx];
        }
        out->amps()[l] = acc;
    }
    return DFSA_OK;
}
}  // extern "C"

// ---- the remaining channels and the expectation value, written against the GLOBAL Choi index (flat = 2^N col + row; bit q = ket bit
// of qubit q, bit q + N = its bra bit) wherever the amplitude lives: an entry that needs other ranks' amplitudes reads them from the
// arena between two barriers over all ranks (every rank makes these calls: the host layer never skips a rank for a channel).
static Amp globalAmp(const dfsa_state* s, uint64_t i) {
    g_remoteAmps += int(i >> s->logNumAmps) != g.rank;
    return s->arr(int(i >> s->logNumAmps), DFSA_AMPS)[i & (s->numAmps - 1)];
}

template <class F>      // f(global index) -> new amplitude, from the OLD global state
static int channelFromGlobal(dfsa_state* s, bool collective, F f) {
    std::vector<Amp> out(s->numAmps);
    if (collective) barrierAll();
    for (uint64_t j = 0; j < s->numAmps; j++) out[j] = f(s->rankShift() | j);
    if (collective) barrierAll();
    memcpy(static_cast<void*>(s->amps()), out.data(), s->numAmps * sizeof(Amp));
    return DFSA_OK;
}
static bool braOnRankBit(const dfsa_state* s, unsigned qb) { return qb + s->numQubits >= s->logNumAmps; }

// rho -> (1 - p) rho + p/3 (X rho X + Y rho Y + Z rho Z): coherences * (1 - 4p/3); populations mixed with 2p/3, 1 - 2p/3
static int simDepol1(dfsa_state* s, unsigned qb, double p, bool collective) {
    const unsigned N = s->numQubits;
    return channelFromGlobal(s, collective, [=](uint64_t i) {
        const Amp a = globalAmp(s, i);
        if (((i >> qb) ^ (i >> (qb + N))) & 1ULL) return (1 - 4 * p / 3) * a;
        return (1 - 2 * p / 3) * a + (2 * p / 3) * globalAmp(s, i ^ (1ULL << qb) ^ (1ULL << (qb + N)));
    });
}
// amplitude damping: rho_00 += p rho_11; rho_11 *= 1 - p; coherences *= sqrt(1 - p)
static int simDamping(dfsa_state* s, unsigned qb, double p, bool collective) {
    const unsigned N = s->numQubits;
    return channelFromGlobal(s, collective, [=](uint64_t i) {
        const Amp a = globalAmp(s, i);
        const unsigned ket = (i >> qb) & 1ULL, bra = (i >> (qb + N)) & 1ULL;
        if (ket != bra) return std::sqrt(1 - p) * a;
        if (ket) return (1 - p) * a;
        return a + p * globalAmp(s, i | (1ULL << qb) | (1ULL << (qb + N)));
    });
}
// keep * rho + (4p/15) I (x) Tr_2 rho on the two qubits; keep = 1 - 16p/15 is the depolarising channel (`corrected`), keep = 1 - 4p/5
// on the populations is what the reference's LOCAL branch computes (local_densitymatrix.hpp:83-108)
static int simDepol2(dfsa_state* s, unsigned q1, unsigned q2, double p, bool corrected, bool collective) {
    const unsigned N = s->numQubits;
    const uint64_t all = (1ULL << q1) | (1ULL << q2) | (1ULL << (q1 + N)) | (1ULL << (q2 + N));
    return channelFromGlobal(s, collective, [=](uint64_t i) {
        const Amp a = globalAmp(s, i);
        const bool same = !((((i >> q1) ^ (i >> (q1 + N))) | ((i >> q2) ^ (i >> (q2 + N)))) & 1ULL);
        if (!same) return (1 - 16 * p / 15) * a;
        Amp sum(0, 0);
        for (unsigned c = 0; c < 4; c++) {
            uint64_t idx = i & ~all;
            if (c & 1u) idx |= (1ULL << q1) | (1ULL << (q1 + N));
            if (c & 2u) idx |= (1ULL << q2) | (1ULL << (q2 + N));
            sum += globalAmp(s, idx);
        }
        return (corrected ? 1 - 16 * p / 15 : 1 - 4 * p / 5) * a + (4 * p / 15) * sum;
    });
}
extern "C" {

int dfsa_k_oneQubitDepolarising(dfsa_state* s, unsigned qb, double prob) {
    touchOp();
    SIM_REQUIRE(s && s->isDensity && qb < s->numQubits && !braOnRankBit(s, qb), "the local kernel needs a qubit whose bra bit is a suffix bit");
    return simDepol1(s, qb, prob, false);
}
int dfsa_xk_depol1Prefix(dfsa_state* s, unsigned qb, unsigned bit, double prob, int pairRank) {
    touchOp();
    SIM_REQUIRE(s && s->isDensity && qb < s->numQubits && braOnRankBit(s, qb), "needs a qubit whose bra bit is a rank bit");
    const unsigned rankBit = qb + s->numQubits - s->logNumAmps;
    SIM_REQUIRE(bit == ((unsigned(g.rank) >> rankBit) & 1u) && pairRank == (g.rank ^ (1 << rankBit)), "bit / pairRank do not belong to this rank and qubit");
    return simDepol1(s, qb, prob, true);
}
int dfsa_k_damping(dfsa_state* s, unsigned qb, double prob) {
    touchOp();
    SIM_REQUIRE(s && s->isDensity && qb < s->numQubits && !braOnRankBit(s, qb), "the local kernel needs a qubit whose bra bit is a suffix bit");
    return simDamping(s, qb, prob, false);
}
int dfsa_xk_dampingPrefix(dfsa_state* s, unsigned qb, unsigned bit, double prob, int pairRank) {
    touchOp();
    SIM_REQUIRE(s && s->isDensity && qb < s->numQubits && braOnRankBit(s, qb), "needs a qubit whose bra bit is a rank bit");
    const unsigned rankBit = qb + s->numQubits - s->logNumAmps;
    SIM_REQUIRE(bit == ((unsigned(g.rank) >> rankBit) & 1u) && pairRank == (g.rank ^ (1 << rankBit)), "bit / pairRank do not belong to this rank and qubit");
    return simDamping(s, qb, prob, true);
}
int dfsa_k_twoQubitDepolarising(dfsa_state* s, unsigned qb1, unsigned qb2, double prob, int corrected) {
    touchOp();
    SIM_REQUIRE(s && s->isDensity && qb1 < s->numQubits && qb2 < s->numQubits && qb1 != qb2 && !braOnRankBit(s, qb1) && !braOnRankBit(s, qb2), "the local kernel needs suffix bra bits");
    return simDepol2(s, qb1, qb2, prob, corrected != 0, false);
}
// the reference's formulas for these two branches (distributed_densitymatrix.hpp:146-237, with their read-after-write) are pinned
// by tests/golden on the GPU; here only the true channel
int dfsa_xk_depol2Pair(dfsa_state* s, unsigned qb1, unsigned qb2, unsigned bit, double prob, int corrected, int pairRank) {
    touchOp();
    SIM_REQUIRE(corrected, "hostsim simulates the corrected channel only on the prefix branches");
    SIM_REQUIRE(s && s->isDensity && qb1 < qb2 && qb2 < s->numQubits && !braOnRankBit(s, qb1) && braOnRankBit(s, qb2), "pair branch: qb1 suffix, qb2 prefix");
    const unsigned rankBit = qb2 + s->numQubits - s->logNumAmps;
    SIM_REQUIRE(bit == ((unsigned(g.rank) >> rankBit) & 1u) && pairRank == (g.rank ^ (1 << rankBit)), "bit / pairRank do not belong to this rank and qubit");
    return simDepol2(s, qb1, qb2, prob, true, true);
}
int dfsa_xk_depol2Quad(dfsa_state* s, unsigned qb1, unsigned qb2, unsigned bit0, unsigned bit1, double prob, int corrected, int pairRank0, int pairRank1) {
    touchOp();
    SIM_REQUIRE(corrected, "hostsim simulates the corrected channel only on the prefix branches");
    SIM_REQUIRE(s && s->isDensity && qb1 < qb2 && qb2 < s->numQubits && braOnRankBit(s, qb1) && braOnRankBit(s, qb2), "quad branch: both bra bits are rank bits");
    const unsigned r0 = qb1 + s->numQubits - s->logNumAmps, r1 = qb2 + s->numQubits - s->logNumAmps;
    SIM_REQUIRE(bit0 == ((unsigned(g.rank) >> r0) & 1u) && bit1 == ((unsigned(g.rank) >> r1) & 1u) && pairRank0 == (g.rank ^ (1 << r0)) && pairRank1 == (g.rank ^ (1 << r1)),
                "bits / pairRanks do not belong to this rank and qubits");
    return simDepol2(s, qb1, qb2, prob, true, true);
}
// this rank's part of sum_t coeff_t Tr(P_t rho) = sum_t coeff_t sum_i <row(i)| P_t |col(i)> rho_flat[i]; row = bra bits, col = ket bits
// (misc.hpp:16-38). The host layer adds the parts up with comm_reduceAmp (*outIsGlobal = 0).
int dfsa_k_expecPauliString(dfsa_state* s, const double* coeffs, unsigned numTerms, const uint32_t* paulis, double out[2]) {
    touchOp();
    SIM_REQUIRE(s && s->isDensity && coeffs && paulis && out, "arguments");
    static const Amp mats[4][2][2] = {{{1, 0}, {0, 1}}, {{0, 1}, {1, 0}}, {{0, Amp(0, -1)}, {Amp(0, 1), 0}}, {{1, 0}, {0, -1}}};
    const unsigned N = s->numQubits;
    Amp total(0, 0);
    for (unsigned t = 0; t < numTerms; t++) {
        Amp termSum(0, 0);
        for (uint64_t j = 0; j < s->numAmps; j++) {
            const uint64_t i = s->rankShift() | j;
            Amp elem(1, 0);
            for (unsigned q = 0; q < N && elem != Amp(0, 0); q++) {
                SIM_REQUIRE(paulis[t * N + q] < 4, "Pauli codes are 0..3");
                elem *= mats[paulis[t * N + q]][(i >> (q + N)) & 1ULL][(i >> q) & 1ULL];
            }
            termSum += elem * s->amps()[j];
        }
        total += coeffs[t] * termSum;
    }
    out[0] = total.real();
    out[1] = total.imag();
    return DFSA_OK;
}
int dfsa_kx_expecPauliString(dfsa_state* s, const double* coeffs, unsigned numTerms, const uint32_t* paulis, double out[2], int* outIsGlobal) {
    if (outIsGlobal) *outIsGlobal = 0;
    return dfsa_k_expecPauliString(s, coeffs, numTerms, paulis, out);
}
int dfsa_k_depol1Combine(dfsa_state*, unsigned, unsigned, double) { setError("hostsim: channel not simulated"); return DFSA_ERR_UNSUPPORTED; }
int dfsa_k_depol2Pair(dfsa_state*, unsigned, unsigned, unsigned, unsigned, double, int) { setError("hostsim: channel not simulated"); return DFSA_ERR_UNSUPPORTED; }
int dfsa_k_depol2Quad(dfsa_state*, unsigned, unsigned, unsigned, unsigned, double, int) { setError("hostsim: channel not simulated"); return DFSA_ERR_UNSUPPORTED; }
int dfsa_k_dampingPrefix(dfsa_state*, unsigned, unsigned, double, int) { setError("hostsim: channel not simulated"); return DFSA_ERR_UNSUPPORTED; }

}  // extern "C"
