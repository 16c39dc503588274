/*
 * dfsa_oracle.c -- CPU restatement (plain C99) of the reference's distributed full-state algorithms.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * as the CHECKER. It is never linked into, called from, or a fallback for the product (the CUDA library
 * fails loudly when it cannot run).
 *
 * What it restates (reference paths relative to /root/reference):
 *   src/bit_maths.hpp (with the setBit bug of :66 fixed), src/local_statevector.hpp,
 *   src/local_densitymatrix.hpp, src/distributed_statevector.hpp, src/distributed_densitymatrix.hpp,
 *   the planners of src/misc.hpp:84-135 and the Pauli element of src/misc.hpp:16-38.
 * The reference is SPMD over MPI ranks; here all P = 2^k "ranks" live in one process as P separate
 * (amps, buffer) array pairs, every algorithm is written as  phase(all ranks) -> exchange -> phase(all ranks),
 * and an exchange is a memcpy between two ranks' arrays (src/communication.hpp:77-113).
 * Indices are 64-bit throughout (the reference silently truncates above 32 index bits).
 *
 * Pinning: tests/test_oracle_golden.py checks every function below against outputs of the real (patched)
 * reference build oracle/_ref/ref_driver at 1/2/4/8 ranks (committed fixtures tests/golden/ npz files, made by
 * tests/golden/make_golden.py) and against an independent dense Kronecker-product ground truth.
 * Known reference quirks are reproduced when orc_set_quirks(1) (default): twoQubitDepolarising's literal
 * formulas and the no-op for Pauli strings without X/Y (local path has 0 inner iterations).
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef double complex amp_t;
typedef uint64_t idx_t;

typedef struct {
    int isDensity;
    int numQubits;            /* n (sv) or N (dm) */
    int numNodes, logNumNodes;
    int logNumAmpsPerNode;
    idx_t numAmpsPerNode;
    amp_t** amps;             /* [rank][local index] */
    amp_t** buffer;
} orc_state;

static int g_quirks = 1;
void orc_set_quirks(int on) { g_quirks = on; }

/* ---------------------------------------------------------------- bit maths (src/bit_maths.hpp:19-125) */

static inline idx_t pow2(int e) { return (idx_t)1 << e; }
static inline int   bit_of(idx_t n, int i) { return (int)((n >> i) & 1); }
static inline idx_t flip_bit(idx_t n, int i) { return n ^ ((idx_t)1 << i); }
static inline idx_t insert_bit(idx_t n, int i, int v) {
    idx_t lo = n & (((idx_t)1 << i) - 1);
    return ((n >> i) << (i + 1)) | ((idx_t)v << i) | lo;
}
/* positions strictly increasing (src/bit_maths.hpp:52-58) */
static inline idx_t insert_bits(idx_t n, const int* pos, int cnt, int v) {
    for (int q = 0; q < cnt; q++) n = insert_bit(n, pos[q], v);
    return n;
}
/* corrected form of src/bit_maths.hpp:62-67 */
static inline idx_t set_bit(idx_t n, int i, int v) { return (n & ~((idx_t)1 << i)) | ((idx_t)v << i); }
/* bit q of `value` goes to position pos[q], caller's order (src/bit_maths.hpp:70-78) */
static inline idx_t set_bits(idx_t n, const int* pos, int cnt, idx_t value) {
    for (int q = 0; q < cnt; q++) n = set_bit(n, pos[q], bit_of(value, q));
    return n;
}
static inline int parity64(idx_t m) { return __builtin_parityll(m); }
static idx_t mask_of(const int* pos, int cnt) { idx_t m = 0; for (int q = 0; q < cnt; q++) m ^= (idx_t)1 << pos[q]; return m; }

static int cmp_int(const void* a, const void* b) { return *(const int*)a - *(const int*)b; }
static void sorted_copy(int* dst, const int* src, int cnt) { memcpy(dst, src, sizeof(int) * cnt); qsort(dst, cnt, sizeof(int), cmp_int); }

/* ---------------------------------------------------------------- state (src/states.hpp:13-69) */

orc_state* orc_create(int isDensity, int numQubits, int numNodes) {
    int k = 0; while ((1 << k) < numNodes) k++;
    if ((1 << k) != numNodes) return NULL;
    if (pow2(numQubits) < (idx_t)numNodes) return NULL;           /* src/states.hpp:35 */
    orc_state* s = (orc_state*)calloc(1, sizeof(orc_state));
    s->isDensity = isDensity; s->numQubits = numQubits; s->numNodes = numNodes; s->logNumNodes = k;
    s->logNumAmpsPerNode = (isDensity ? 2 * numQubits : numQubits) - k;
    s->numAmpsPerNode = pow2(s->logNumAmpsPerNode);
    s->amps = (amp_t**)calloc(numNodes, sizeof(amp_t*));
    s->buffer = (amp_t**)calloc(numNodes, sizeof(amp_t*));
    for (int r = 0; r < numNodes; r++) {
        s->amps[r] = (amp_t*)calloc(s->numAmpsPerNode, sizeof(amp_t));
        s->buffer[r] = (amp_t*)calloc(s->numAmpsPerNode, sizeof(amp_t));
    }
    return s;
}

void orc_destroy(orc_state* s) {
    if (!s) return;
    for (int r = 0; r < s->numNodes; r++) { free(s->amps[r]); free(s->buffer[r]); }
    free(s->amps); free(s->buffer); free(s);
}

int orc_log_amps_per_node(const orc_state* s) { return s->logNumAmpsPerNode; }
int orc_num_qubits(const orc_state* s) { return s->numQubits; }

/* global array, rank r owns [r*A, (r+1)*A)  (src/states.hpp:41-44) */
void orc_set_amps(orc_state* s, const double* interleaved) {
    for (int r = 0; r < s->numNodes; r++)
        memcpy(s->amps[r], interleaved + 2 * (idx_t)r * s->numAmpsPerNode, sizeof(amp_t) * s->numAmpsPerNode);
}
void orc_get_amps(const orc_state* s, double* interleaved) {
    for (int r = 0; r < s->numNodes; r++)
        memcpy(interleaved + 2 * (idx_t)r * s->numAmpsPerNode, s->amps[r], sizeof(amp_t) * s->numAmpsPerNode);
}

/* the synthetic state of SURVEY 8(d): counter-based hash of the global index */
static uint64_t splitmix64(uint64_t seed, uint64_t k) {
    uint64_t z = seed + (k + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static double hash_real(uint64_t seed, uint64_t k) { return (double)(splitmix64(seed, k) >> 11) * (1.0 / 9007199254740992.0) - 0.5; }
void orc_init_hash(orc_state* s, uint64_t seed) {
    for (int r = 0; r < s->numNodes; r++) {
        idx_t first = (idx_t)r * s->numAmpsPerNode;
        #pragma omp parallel for
        for (idx_t j = 0; j < s->numAmpsPerNode; j++)
            s->amps[r][j] = hash_real(seed, 2 * (first + j)) + I * hash_real(seed, 2 * (first + j) + 1);
    }
}

/* pairwise exchange: every rank r in `who` sends send[r][so..so+n) to pair(r)'s recv[..][ro..ro+n)
 * (src/communication.hpp:77-113) */
static void exchange(orc_state* s, amp_t** send, idx_t so, amp_t** recv, idx_t ro, idx_t n, const int* pairOf) {
    for (int r = 0; r < s->numNodes; r++)
        if (pairOf[r] >= 0 && pairOf[r] != r)
            memcpy(recv[pairOf[r]] + ro, send[r] + so, sizeof(amp_t) * n);
}

/* gate arguments arrive as row-major interleaved (re,im) doubles */
#define G(g, dim, r, c) ((g)[2 * ((idx_t)(r) * (dim) + (c))] + I * (g)[2 * ((idx_t)(r) * (dim) + (c)) + 1])

/* ---------------------------------------------------------------- local state-vector loops */

/* src/local_statevector.hpp:14-29 */
static void local_oneTarg(orc_state* s, int r, int target, const double* gate) {
    amp_t g00 = G(gate, 2, 0, 0), g01 = G(gate, 2, 0, 1), g10 = G(gate, 2, 1, 0), g11 = G(gate, 2, 1, 1);
    amp_t* a = s->amps[r];
    idx_t its = s->numAmpsPerNode / 2;
    #pragma omp parallel for
    for (idx_t j = 0; j < its; j++) {
        idx_t i0 = insert_bit(j, target, 0), i1 = flip_bit(i0, target);
        amp_t a0 = a[i0], a1 = a[i1];
        a[i0] = g00 * a0 + g01 * a1;
        a[i1] = g10 * a0 + g11 * a1;
    }
}

/* src/local_statevector.hpp:32-51 */
static void local_manyCtrlOneTarg(orc_state* s, int r, const int* ctrls, int nc, int target, const double* gate) {
    amp_t g00 = G(gate, 2, 0, 0), g01 = G(gate, 2, 0, 1), g10 = G(gate, 2, 1, 0), g11 = G(gate, 2, 1, 1);
    int qs[64]; memcpy(qs, ctrls, sizeof(int) * nc); qs[nc] = target; qsort(qs, nc + 1, sizeof(int), cmp_int);
    amp_t* a = s->amps[r];
    idx_t its = s->numAmpsPerNode >> (nc + 1);
    #pragma omp parallel for
    for (idx_t j = 0; j < its; j++) {
        idx_t i1 = insert_bits(j, qs, nc + 1, 1), i0 = flip_bit(i1, target);
        amp_t a0 = a[i0], a1 = a[i1];
        a[i0] = g00 * a0 + g01 * a1;
        a[i1] = g10 * a0 + g11 * a1;
    }
}

/* src/local_statevector.hpp:54-69 */
static void local_swap(orc_state* s, int r, int q1, int q2) {
    if (q1 > q2) { int t = q1; q1 = q2; q2 = t; }
    amp_t* a = s->amps[r];
    idx_t its = s->numAmpsPerNode / 4;
    #pragma omp parallel for
    for (idx_t k = 0; k < its; k++) {
        idx_t j11 = insert_bit(insert_bit(k, q1, 1), q2, 1);
        idx_t j10 = flip_bit(j11, q1), j01 = flip_bit(j11, q2);
        amp_t t = a[j01]; a[j01] = a[j10]; a[j10] = t;
    }
}

/* src/local_statevector.hpp:72-99 : gate bit i <-> targets[i] (caller order); zero insertion on sorted copy */
static void local_manyTarg(orc_state* s, int r, const int* targets, int nt, const double* gate) {
    idx_t dim = pow2(nt);
    int sorted[64]; sorted_copy(sorted, targets, nt);
    amp_t* a = s->amps[r];
    idx_t outer = s->numAmpsPerNode / dim;
    #pragma omp parallel
    {
        amp_t* cache = (amp_t*)malloc(sizeof(amp_t) * dim);
        #pragma omp for
        for (idx_t k = 0; k < outer; k++) {
            idx_t base = insert_bits(k, sorted, nt, 0);
            for (idx_t j = 0; j < dim; j++) cache[j] = a[set_bits(base, targets, nt, j)];
            for (idx_t j = 0; j < dim; j++) {
                amp_t acc = 0;
                for (idx_t l = 0; l < dim; l++) acc += G(gate, dim, j, l) * cache[l];
                a[set_bits(base, targets, nt, j)] = acc;
            }
        }
        free(cache);
    }
}

/* src/local_statevector.hpp:102-135 */
static void local_pauli(orc_state* s, int r, const int* suffixXY, int nxy, amp_t powI, idx_t maskXY, idx_t maskYZ, amp_t thisFac, amp_t otherFac) {
    int sorted[64]; sorted_copy(sorted, suffixXY, nxy);
    amp_t* a = s->amps[r];
    idx_t outer = s->numAmpsPerNode >> nxy;
    idx_t inner = pow2(nxy) / 2;        /* == 0 when nxy == 0: the reference's all-Z no-op */
    idx_t rankShift = (idx_t)r << s->logNumAmpsPerNode;
    if (nxy == 0 && !g_quirks) {
        /* what the operator actually is when no X/Y is present: a diagonal of signs */
        #pragma omp parallel for
        for (idx_t j = 0; j < s->numAmpsPerNode; j++) {
            amp_t b = (1. - 2. * parity64((rankShift | j) & maskYZ)) * powI;
            a[j] = thisFac * a[j] + otherFac * b * a[j];
        }
        return;
    }
    #pragma omp parallel for
    for (idx_t k = 0; k < outer; k++) {
        idx_t h = insert_bits(k, sorted, nxy, 0);
        for (idx_t l = 0; l < inner; l++) {
            idx_t j0 = set_bits(h, sorted, nxy, l), j1 = j0 ^ maskXY;
            amp_t b0 = (1. - 2. * parity64((rankShift | j0) & maskYZ)) * powI;
            amp_t b1 = (1. - 2. * parity64((rankShift | j1) & maskYZ)) * powI;
            amp_t a0 = a[j0], a1 = a[j1];
            a[j0] = thisFac * a0 + otherFac * b1 * a1;
            a[j1] = thisFac * a1 + otherFac * b0 * a0;
        }
    }
}

/* src/local_statevector.hpp:138-153 */
static void local_phaseGadget(orc_state* s, int r, const int* targets, int nt, double theta) {
    idx_t rankShift = (idx_t)r << s->logNumAmpsPerNode, mask = mask_of(targets, nt);
    amp_t facs[2] = { cos(theta) + I * sin(theta), cos(theta) - I * sin(theta) };
    amp_t* a = s->amps[r];
    #pragma omp parallel for
    for (idx_t j = 0; j < s->numAmpsPerNode; j++)
        a[j] *= facs[parity64((rankShift | j) & mask)];
}

/* ---------------------------------------------------------------- distributed state-vector API */

/* src/distributed_statevector.hpp:18-40 */
void orc_sv_oneTargGate(orc_state* s, int target, const double* gate) {
    int L = s->logNumAmpsPerNode;
    if (target < L) { for (int r = 0; r < s->numNodes; r++) local_oneTarg(s, r, target, gate); return; }
    int rt = target - L;
    int pairOf[64 * 4];
    for (int r = 0; r < s->numNodes; r++) pairOf[r] = (int)flip_bit(r, rt);
    exchange(s, s->amps, 0, s->buffer, 0, s->numAmpsPerNode, pairOf);
    for (int r = 0; r < s->numNodes; r++) {
        int b = bit_of(r, rt);
        amp_t f0 = G(gate, 2, b, b), f1 = G(gate, 2, b, !b);
        amp_t* a = s->amps[r]; const amp_t* buf = s->buffer[r];
        #pragma omp parallel for
        for (idx_t i = 0; i < s->numAmpsPerNode; i++) a[i] = f0 * a[i] + f1 * buf[i];
    }
}

/* src/distributed_statevector.hpp:43-106 */
void orc_sv_manyCtrlOneTargGate(orc_state* s, const int* ctrls, int nc, int target, const double* gate) {
    int L = s->logNumAmpsPerNode;
    int prefix[64], suffix[64], np = 0, ns = 0;
    for (int q = 0; q < nc; q++) { if (ctrls[q] >= L) prefix[np++] = ctrls[q] - L; else suffix[ns++] = ctrls[q]; }
    idx_t prefixMask = mask_of(prefix, np);
    int active[256];
    for (int r = 0; r < s->numNodes; r++) active[r] = (((idx_t)r & prefixMask) == prefixMask);   /* :92-93 */

    if (target < L) {
        for (int r = 0; r < s->numNodes; r++) if (active[r]) local_manyCtrlOneTarg(s, r, suffix, ns, target, gate);
        return;
    }
    int rt = target - L;
    int pairOf[256];
    for (int r = 0; r < s->numNodes; r++) pairOf[r] = active[r] ? (int)flip_bit(r, rt) : -1;

    if (ns == 0) {                                  /* :100-101 degrade to the uncontrolled prefix gate */
        exchange(s, s->amps, 0, s->buffer, 0, s->numAmpsPerNode, pairOf);
        for (int r = 0; r < s->numNodes; r++) if (active[r]) {
            int b = bit_of(r, rt);
            amp_t f0 = G(gate, 2, b, b), f1 = G(gate, 2, b, !b);
            amp_t* a = s->amps[r]; const amp_t* buf = s->buffer[r];
            #pragma omp parallel for
            for (idx_t i = 0; i < s->numAmpsPerNode; i++) a[i] = f0 * a[i] + f1 * buf[i];
        }
        return;
    }
    /* :43-78 pack the ctrl=1 subset, swap sub-buffers, combine */
    int sortedCtrls[64]; sorted_copy(sortedCtrls, suffix, ns);
    idx_t m = s->numAmpsPerNode >> ns;
    for (int r = 0; r < s->numNodes; r++) if (active[r]) {
        #pragma omp parallel for
        for (idx_t j = 0; j < m; j++) s->buffer[r][j] = s->amps[r][insert_bits(j, sortedCtrls, ns, 1)];
    }
    exchange(s, s->buffer, 0, s->buffer, m, m, pairOf);
    for (int r = 0; r < s->numNodes; r++) if (active[r]) {
        int b = bit_of(r, rt);
        amp_t f0 = G(gate, 2, b, b), f1 = G(gate, 2, b, !b);
        #pragma omp parallel for
        for (idx_t j = 0; j < m; j++) {
            idx_t k = insert_bits(j, sortedCtrls, ns, 1);
            s->amps[r][k] = f0 * s->amps[r][k] + f1 * s->buffer[r][j + m];
        }
    }
}

/* src/distributed_statevector.hpp:109-187 */
void orc_sv_swapGate(orc_state* s, int q1, int q2) {
    if (q1 > q2) { int t = q1; q1 = q2; q2 = t; }
    int L = s->logNumAmpsPerNode;
    idx_t A = s->numAmpsPerNode;
    int pairOf[256];
    if (q2 < L) { for (int r = 0; r < s->numNodes; r++) local_swap(s, r, q1, q2); return; }

    if (q1 >= L) {                                   /* both prefix: half the ranks trade whole shards (:120-137) */
        int a1 = q1 - L, a2 = q2 - L;
        for (int r = 0; r < s->numNodes; r++)
            pairOf[r] = (bit_of(r, a1) != bit_of(r, a2)) ? (int)flip_bit(flip_bit(r, a1), a2) : -1;
        exchange(s, s->amps, 0, s->buffer, 0, A, pairOf);
        for (int r = 0; r < s->numNodes; r++) if (pairOf[r] >= 0) memcpy(s->amps[r], s->buffer[r], sizeof(amp_t) * A);
        return;
    }
    int a2 = q2 - L;
    idx_t half = A / 2;
    for (int r = 0; r < s->numNodes; r++) pairOf[r] = (int)flip_bit(r, a2);

    if (q1 == L - 1) {                               /* contiguous half (:140-157) */
        /* the send offset depends on the sender's rank bit, so exchange rank by rank */
        for (int r = 0; r < s->numNodes; r++) {
            idx_t off = half * (idx_t)(!bit_of(r, a2));
            memcpy(s->buffer[pairOf[r]], s->amps[r] + off, sizeof(amp_t) * half);
        }
        for (int r = 0; r < s->numNodes; r++) {
            idx_t off = half * (idx_t)(!bit_of(r, a2));
            memcpy(s->amps[r] + off, s->buffer[r], sizeof(amp_t) * half);
        }
        return;
    }
    /* packed half (:160-186) */
    for (int r = 0; r < s->numNodes; r++) {
        int b1 = !bit_of(r, a2);
        #pragma omp parallel for
        for (idx_t k = 0; k < half; k++) s->buffer[r][k] = s->amps[r][insert_bit(k, q1, b1)];
    }
    exchange(s, s->buffer, 0, s->buffer, half, half, pairOf);
    for (int r = 0; r < s->numNodes; r++) {
        int b1 = !bit_of(r, a2);
        #pragma omp parallel for
        for (idx_t k = 0; k < half; k++) s->amps[r][insert_bit(k, q1, b1)] = s->buffer[r][k + half];
    }
}

/* planner of src/distributed_statevector.hpp:193-210: prefix targets (caller order) take the lowest free suffix qubits */
void orc_plan_manyTarg(int logNumAmpsPerNode, const int* targets, int nt, int* newTargs) {
    idx_t mask = mask_of(targets, nt);
    int minFree = 0;
    while (bit_of(mask, minFree)) minFree++;
    for (int q = 0; q < nt; q++) {
        if (targets[q] < logNumAmpsPerNode) newTargs[q] = targets[q];
        else {
            newTargs[q] = minFree++;
            while (bit_of(mask, minFree)) minFree++;
        }
    }
}

/* src/distributed_statevector.hpp:190-224 */
void orc_sv_manyTargGate(orc_state* s, const int* targets, int nt, const double* gate) {
    int newTargs[64];
    orc_plan_manyTarg(s->logNumAmpsPerNode, targets, nt, newTargs);
    for (int q = 0; q < nt; q++) if (newTargs[q] != targets[q]) orc_sv_swapGate(s, newTargs[q], targets[q]);
    for (int r = 0; r < s->numNodes; r++) local_manyTarg(s, r, newTargs, nt, gate);
    for (int q = 0; q < nt; q++) if (newTargs[q] != targets[q]) orc_sv_swapGate(s, newTargs[q], targets[q]);
}

/* src/distributed_statevector.hpp:227-276 */
static void pauli_tensor_or_gadget(orc_state* s, const int* targets, const int* paulis, int nt, amp_t thisFac, amp_t otherFac) {
    int L = s->logNumAmpsPerNode;
    amp_t powI = 1;
    for (int q = 0; q < nt; q++) if (paulis[q] == 2) powI *= I;
    idx_t prefixFlip = 0, maskXY = 0, maskYZ = 0;
    int suffixXY[64], nxy = 0;
    for (int q = 0; q < nt; q++) {
        int isXY = (paulis[q] == 1 || paulis[q] == 2), isYZ = (paulis[q] == 2 || paulis[q] == 3);
        if (isXY) { if (targets[q] >= L) prefixFlip ^= (idx_t)1 << (targets[q] - L); else { suffixXY[nxy++] = targets[q]; maskXY ^= (idx_t)1 << targets[q]; } }
        if (isYZ) maskYZ ^= (idx_t)1 << targets[q];
    }
    if (prefixFlip == 0) {
        for (int r = 0; r < s->numNodes; r++) local_pauli(s, r, suffixXY, nxy, powI, maskXY, maskYZ, thisFac, otherFac);
        return;
    }
    int pairOf[256];
    for (int r = 0; r < s->numNodes; r++) pairOf[r] = (int)((idx_t)r ^ prefixFlip);
    exchange(s, s->amps, 0, s->buffer, 0, s->numAmpsPerNode, pairOf);
    for (int r = 0; r < s->numNodes; r++) {              /* :227-241, sign from the PARTNER's global index */
        idx_t rankShift = (idx_t)pairOf[r] << L;
        amp_t* a = s->amps[r]; const amp_t* buf = s->buffer[r];
        #pragma omp parallel for
        for (idx_t j0 = 0; j0 < s->numAmpsPerNode; j0++) {
            idx_t j1 = j0 ^ maskXY;
            amp_t b1 = (1. - 2. * parity64((rankShift | j1) & maskYZ)) * powI;
            a[j0] = thisFac * a[j0] + otherFac * b1 * buf[j1];
        }
    }
}

/* src/distributed_statevector.hpp:279-284 */
void orc_sv_pauliTensor(orc_state* s, const int* targets, const int* paulis, int nt) { pauli_tensor_or_gadget(s, targets, paulis, nt, 0., 1.); }
/* src/distributed_statevector.hpp:287-292 */
void orc_sv_pauliGadget(orc_state* s, const int* targets, const int* paulis, int nt, double theta) { pauli_tensor_or_gadget(s, targets, paulis, nt, cos(theta), I * sin(theta)); }
/* src/distributed_statevector.hpp:295-298 */
void orc_sv_phaseGadget(orc_state* s, const int* targets, int nt, double theta) { for (int r = 0; r < s->numNodes; r++) local_phaseGadget(s, r, targets, nt, theta); }

/* ---------------------------------------------------------------- density-matrix API */

static void shifted(int* dst, const int* src, int n, int by) { for (int q = 0; q < n; q++) dst[q] = src[q] + by; }

/* src/distributed_densitymatrix.hpp:15-25 */
void orc_dm_manyTargGate(orc_state* s, const int* targets, int nt, const double* gate) {
    idx_t dim = pow2(nt);
    orc_sv_manyTargGate(s, targets, nt, gate);
    int bra[64]; shifted(bra, targets, nt, s->numQubits);
    double* conjGate = (double*)malloc(sizeof(double) * 2 * dim * dim);
    for (idx_t e = 0; e < dim * dim; e++) { conjGate[2 * e] = gate[2 * e]; conjGate[2 * e + 1] = -gate[2 * e + 1]; }
    orc_sv_manyTargGate(s, bra, nt, conjGate);
    free(conjGate);
}
/* src/distributed_densitymatrix.hpp:28-33 */
void orc_dm_swapGate(orc_state* s, int q1, int q2) { orc_sv_swapGate(s, q1, q2); orc_sv_swapGate(s, q1 + s->numQubits, q2 + s->numQubits); }

static int odd_num_y(const int* paulis, int nt) { int odd = 0; for (int q = 0; q < nt; q++) odd ^= (paulis[q] == 2); return odd; }

/* src/distributed_densitymatrix.hpp:36-50 */
void orc_dm_pauliTensor(orc_state* s, const int* targets, const int* paulis, int nt) {
    int bra[64]; shifted(bra, targets, nt, s->numQubits);
    orc_sv_pauliTensor(s, targets, paulis, nt);
    orc_sv_pauliTensor(s, bra, paulis, nt);
    if (odd_num_y(paulis, nt))
        for (int r = 0; r < s->numNodes; r++) {
            #pragma omp parallel for
            for (idx_t j = 0; j < s->numAmpsPerNode; j++) s->amps[r][j] *= -1;
        }
}
/* src/distributed_densitymatrix.hpp:53-64 */
void orc_dm_pauliGadget(orc_state* s, const int* targets, const int* paulis, int nt, double theta) {
    int bra[64]; shifted(bra, targets, nt, s->numQubits);
    orc_sv_pauliGadget(s, targets, paulis, nt, theta);
    orc_sv_pauliGadget(s, bra, paulis, nt, odd_num_y(paulis, nt) ? theta : -theta);
}
/* src/distributed_densitymatrix.hpp:67-76 */
void orc_dm_phaseGadget(orc_state* s, const int* targets, int nt, double theta) {
    int bra[64]; shifted(bra, targets, nt, s->numQubits);
    orc_sv_phaseGadget(s, targets, nt, theta);
    orc_sv_phaseGadget(s, bra, nt, -theta);
}

/* src/misc.hpp:58-81 : superOp = sum_K conj(K) (x) K, row r = i*d + k, col c = j*d + l */
void orc_superoperator(const double* krausOps, int numOps, int nt, double* superOp) {
    idx_t d = pow2(nt), D = d * d;
    memset(superOp, 0, sizeof(double) * 2 * D * D);
    for (int o = 0; o < numOps; o++) {
        const double* K = krausOps + (idx_t)o * 2 * d * d;
        for (idx_t i = 0; i < d; i++) for (idx_t j = 0; j < d; j++) for (idx_t k = 0; k < d; k++) for (idx_t l = 0; l < d; l++) {
            amp_t term = conj(G(K, d, i, j)) * G(K, d, k, l);
            idx_t e = (i * d + k) * D + (j * d + l);
            superOp[2 * e] += creal(term); superOp[2 * e + 1] += cimag(term);
        }
    }
}
/* src/distributed_densitymatrix.hpp:79-89 */
void orc_dm_krausMap(orc_state* s, const double* krausOps, int numOps, const int* targets, int nt) {
    idx_t D = pow2(2 * nt);
    double* superOp = (double*)malloc(sizeof(double) * 2 * D * D);
    orc_superoperator(krausOps, numOps, nt, superOp);
    int ext[64]; memcpy(ext, targets, sizeof(int) * nt); shifted(ext + nt, targets, nt, s->numQubits);
    orc_sv_manyTargGate(s, ext, 2 * nt, superOp);
    free(superOp);
}

/* src/local_densitymatrix.hpp:12-42 (via src/distributed_densitymatrix.hpp:92) */
void orc_dm_oneQubitDephasing(orc_state* s, int qb, double prob) {
    amp_t fac = 1 - 2 * prob;
    int N = s->numQubits, thr = N - s->logNumNodes;
    for (int r = 0; r < s->numNodes; r++) {
        amp_t* a = s->amps[r];
        if (qb >= thr) {
            int b = !bit_of(r, qb - thr);
            idx_t its = s->numAmpsPerNode / 2;
            #pragma omp parallel for
            for (idx_t k = 0; k < its; k++) a[insert_bit(k, qb, b)] *= fac;
        } else {
            idx_t its = s->numAmpsPerNode / 4;
            #pragma omp parallel for
            for (idx_t k = 0; k < its; k++) {
                a[insert_bit(insert_bit(k, qb, 1), qb + N, 0)] *= fac;
                a[insert_bit(insert_bit(k, qb, 0), qb + N, 1)] *= fac;
            }
        }
    }
}

/* src/local_densitymatrix.hpp:45-60 (via :98) */
void orc_dm_twoQubitDephasing(orc_state* s, int q1, int q2, double prob) {
    int N = s->numQubits;
    amp_t term = -4 * prob / 3;
    for (int r = 0; r < s->numNodes; r++) {
        idx_t rankShift = (idx_t)r << s->logNumAmpsPerNode;
        amp_t* a = s->amps[r];
        #pragma omp parallel for
        for (idx_t j = 0; j < s->numAmpsPerNode; j++) {
            idx_t i = rankShift | j;
            int b1 = bit_of(i, q1) ^ bit_of(i, q1 + N), b2 = bit_of(i, q2) ^ bit_of(i, q2 + N);
            amp_t flag = (double)(b1 | b2);
            a[j] *= flag * term + 1.;
        }
    }
}

/* src/distributed_densitymatrix.hpp:104-143 + src/local_densitymatrix.hpp:63-81 */
void orc_dm_oneQubitDepolarising(orc_state* s, int qb, double prob) {
    amp_t c1 = 2 * prob / 3, c2 = 1 - 2 * prob / 3, c3 = 1 - 4 * prob / 3;
    int N = s->numQubits, thr = N - s->logNumNodes;
    if (qb < thr) {
        for (int r = 0; r < s->numNodes; r++) {
            amp_t* a = s->amps[r];
            idx_t its = s->numAmpsPerNode / 4;
            #pragma omp parallel for
            for (idx_t k = 0; k < its; k++) {
                idx_t j00 = insert_bit(insert_bit(k, qb, 0), qb + N, 0);
                idx_t j01 = flip_bit(j00, qb), j10 = flip_bit(j00, qb + N), j11 = flip_bit(j01, qb + N);
                amp_t a00 = a[j00];
                a[j00] = c2 * a00 + c1 * a[j11];
                a[j01] *= c3;
                a[j10] *= c3;
                a[j11] = c1 * a00 + c2 * a[j11];
            }
        }
        return;
    }
    int sh = qb - thr;
    idx_t half = s->numAmpsPerNode / 2;
    int pairOf[256];
    for (int r = 0; r < s->numNodes; r++) {
        int b = bit_of(r, sh);
        pairOf[r] = (int)flip_bit(r, sh);
        #pragma omp parallel for
        for (idx_t k = 0; k < half; k++) s->buffer[r][k] = s->amps[r][insert_bit(k, qb, b)];
    }
    exchange(s, s->buffer, 0, s->buffer, half, half, pairOf);
    for (int r = 0; r < s->numNodes; r++) {
        int b = bit_of(r, sh);
        amp_t* a = s->amps[r];
        #pragma omp parallel for
        for (idx_t k = 0; k < half; k++) a[insert_bit(k, qb, !b)] *= c3;
        #pragma omp parallel for
        for (idx_t k = 0; k < half; k++) { idx_t j = insert_bit(k, qb, b); a[j] = c2 * a[j] + c1 * s->buffer[r][k + half]; }
    }
}

/* src/distributed_densitymatrix.hpp:241-263 with its three branches (:146, :187, src/local_densitymatrix.hpp:83).
 * The formulas are reproduced literally, including the read-after-write of :181->:182 and the overwrite of :236
 * (this is not the depolarising channel; parity with the reference is what is restated here). */
void orc_dm_twoQubitDepolarising(orc_state* s, int qb1, int qb2, double prob) {
    if (qb1 > qb2) { int t = qb1; qb1 = qb2; qb2 = t; }
    amp_t c1 = 1 - 4 * prob / 5, c2 = 4 * prob / 15, c3 = -16 * prob / 15;
    int N = s->numQubits, thr = N - s->logNumNodes;
    idx_t A = s->numAmpsPerNode;
    int pairOf[256];

    if (qb2 < thr) {                                  /* src/local_densitymatrix.hpp:83-108 */
        int q0 = qb1, q1 = qb2, q2 = qb1 + N, q3 = qb2 + N;
        for (int r = 0; r < s->numNodes; r++) {
            amp_t* a = s->amps[r];
            #pragma omp parallel for
            for (idx_t j = 0; j < A; j++) {
                int f1 = !(bit_of(j, q0) ^ bit_of(j, q2)), f2 = !(bit_of(j, q1) ^ bit_of(j, q3));
                a[j] *= 1. + c3 * (double)(!(f1 & f2));
            }
            #pragma omp parallel for
            for (idx_t k = 0; k < A / 16; k++) {
                idx_t j0000 = insert_bit(insert_bit(insert_bit(insert_bit(k, q0, 0), q1, 0), q2, 0), q3, 0);
                idx_t j0101 = flip_bit(flip_bit(j0000, q2), q0);
                idx_t j1010 = flip_bit(flip_bit(j0000, q3), q1);
                idx_t j1111 = flip_bit(flip_bit(j0101, q3), q1);
                amp_t term = a[j0000] + a[j0101] + a[j1010] + a[j1111];
                a[j0000] = c1 * a[j0000] + c2 * term;
                a[j0101] = c1 * a[j0101] + c2 * term;
                a[j1010] = c1 * a[j1010] + c2 * term;
                a[j1111] = c1 * a[j1111] + c2 * term;
            }
        }
        return;
    }
    if (qb1 < thr) {                                  /* pair: :146-184 */
        int q0 = qb1, q1 = qb2, q2 = qb1 + N, alt1 = qb2 - thr;
        idx_t eighth = A / 8;
        for (int r = 0; r < s->numNodes; r++) {
            int b = bit_of(r, alt1);
            amp_t* a = s->amps[r];
            pairOf[r] = (int)flip_bit(r, alt1);
            #pragma omp parallel for
            for (idx_t j = 0; j < A; j++) {
                int f1 = bit_of(j, q0) == bit_of(j, q2), f2 = bit_of(j, q1) == b;
                a[j] *= 1. + c3 * (double)(!(f1 & f2));
            }
            #pragma omp parallel for
            for (idx_t k = 0; k < eighth; k++) {
                idx_t j000 = insert_bit(insert_bit(insert_bit(k, q0, 0), q1, 0), q2, 0);
                idx_t j0b0 = set_bit(j000, q1, b), j1b1 = flip_bit(flip_bit(j0b0, q2), q0);
                s->buffer[r][k] = a[j0b0] + a[j1b1];
            }
        }
        exchange(s, s->buffer, 0, s->buffer, eighth, eighth, pairOf);
        for (int r = 0; r < s->numNodes; r++) {
            int b = bit_of(r, alt1);
            amp_t* a = s->amps[r];
            #pragma omp parallel for
            for (idx_t k = 0; k < eighth; k++) {
                idx_t j000 = insert_bit(insert_bit(insert_bit(k, q0, 0), q1, 0), q2, 0);
                idx_t j0b0 = set_bit(j000, q1, b), j1b1 = flip_bit(flip_bit(j0b0, q2), q0);
                amp_t recv = s->buffer[r][k + eighth];
                a[j0b0] = c1 * a[j0b0] + c2 * (a[j1b1] + recv);
                a[j1b1] = c1 * a[j1b1] + c2 * (a[j0b0] + recv);      /* reads the value written one line up */
            }
        }
        return;
    }
    {                                                 /* quad: :187-238 */
        int q0 = qb1, q1 = qb2, alt0 = qb1 - thr, alt1 = qb2 - thr;
        idx_t quarter = A / 4;
        for (int r = 0; r < s->numNodes; r++) {
            int b0 = bit_of(r, alt0), b1 = bit_of(r, alt1);
            amp_t* a = s->amps[r];
            #pragma omp parallel for
            for (idx_t j = 0; j < A; j++) {
                int f1 = bit_of(j, q0) == b0, f2 = bit_of(j, q1) == b1;
                a[j] *= 1. + c3 * (double)(!(f1 & f2));
            }
            #pragma omp parallel for
            for (idx_t k = 0; k < quarter; k++) s->buffer[r][k] = a[insert_bit(insert_bit(k, q0, b0), q1, b1)];
            pairOf[r] = (int)flip_bit(r, alt0);
        }
        exchange(s, s->buffer, 0, s->buffer, quarter, quarter, pairOf);
        for (int r = 0; r < s->numNodes; r++) {
            int b0 = bit_of(r, alt0), b1 = bit_of(r, alt1);
            amp_t* a = s->amps[r];
            #pragma omp parallel for
            for (idx_t k = 0; k < quarter; k++) {
                idx_t j = insert_bit(insert_bit(k, q0, b0), q1, b1);
                amp_t v = c1 * a[j] + c2 * s->buffer[r][k + quarter];
                a[j] = v; s->buffer[r][k] = v;
            }
            pairOf[r] = (int)flip_bit(r, alt1);
        }
        exchange(s, s->buffer, 0, s->buffer, quarter, quarter, pairOf);
        amp_t c4 = c2 / c1;
        for (int r = 0; r < s->numNodes; r++) {
            int b0 = bit_of(r, alt0), b1 = bit_of(r, alt1);
            amp_t* a = s->amps[r];
            #pragma omp parallel for
            for (idx_t k = 0; k < quarter; k++) a[insert_bit(insert_bit(k, q0, b0), q1, b1)] = c4 * s->buffer[r][k + quarter];
        }
    }
}

/* src/distributed_densitymatrix.hpp:266-319 + src/local_densitymatrix.hpp:111-131 */
void orc_dm_damping(orc_state* s, int qb, double prob) {
    int N = s->numQubits, thr = N - s->logNumNodes;
    amp_t c1 = sqrt(1 - prob), c2 = 1 - prob;
    if (qb < thr) {
        for (int r = 0; r < s->numNodes; r++) {
            amp_t* a = s->amps[r];
            #pragma omp parallel for
            for (idx_t k = 0; k < s->numAmpsPerNode / 4; k++) {
                idx_t j00 = insert_bit(insert_bit(k, qb, 0), qb + N, 0);
                idx_t j01 = flip_bit(j00, qb), j10 = flip_bit(j00, qb + N), j11 = flip_bit(j01, qb + N);
                a[j00] += prob * a[j11];
                a[j01] *= c1;
                a[j10] *= c1;
                a[j11] *= c2;
            }
        }
        return;
    }
    idx_t half = s->numAmpsPerNode / 2;
    int sh = qb - thr;
    for (int r = 0; r < s->numNodes; r++) if (bit_of(r, sh) == 1) {      /* senders: pack + scale, one-way send */
        amp_t* a = s->amps[r];
        #pragma omp parallel for
        for (idx_t k = 0; k < half; k++) { idx_t j = insert_bit(k, qb, 1); s->buffer[r][k] = a[j]; a[j] *= c2; }
        memcpy(s->buffer[flip_bit(r, sh)], s->buffer[r], sizeof(amp_t) * half);
    }
    for (int r = 0; r < s->numNodes; r++) {
        int b = bit_of(r, sh);
        amp_t* a = s->amps[r];
        #pragma omp parallel for
        for (idx_t k = 0; k < half; k++) a[insert_bit(k, qb, !b)] *= c1;
        if (b == 0) {
            #pragma omp parallel for
            for (idx_t k = 0; k < half; k++) a[insert_bit(k, qb, 0)] += prob * s->buffer[r][k];
        }
    }
}

/* src/misc.hpp:16-38 evaluated literally: product over qubits of P_q[rowBit][colBit] */
static amp_t pauli_elem(const int* codes, int N, idx_t flat) {
    amp_t elem = 1.;
    for (int q = 0; q < N; q++) {
        int col = bit_of(flat, q), row = bit_of(flat, q + N);
        amp_t f;
        switch (codes[q]) {
            case 0: f = (row == col) ? 1. : 0.; break;
            case 1: f = (row != col) ? 1. : 0.; break;
            case 2: f = (row == col) ? 0. : (row ? I : -I); break;
            default: f = (row != col) ? 0. : (row ? -1. : 1.); break;
        }
        elem *= f;
    }
    return elem;
}

/* src/distributed_densitymatrix.hpp:322-344 ; out = (re, im) */
void orc_dm_expecPauliString(orc_state* s, const double* coeffs, int numTerms, const int* paulis, double* out) {
    int N = s->numQubits;
    amp_t total = 0;
    for (int r = 0; r < s->numNodes; r++) {
        double re = 0, im = 0;
        #pragma omp parallel for reduction(+:re,im)
        for (idx_t j = 0; j < s->numAmpsPerNode; j++) {
            idx_t i = ((idx_t)r << s->logNumAmpsPerNode) | j;
            amp_t term = 0;
            for (int t = 0; t < numTerms; t++) term += pauli_elem(paulis + (idx_t)t * N, N, i) * coeffs[t];
            amp_t v = term * s->amps[r][j];
            re += creal(v); im += cimag(v);
        }
        total += re + I * im;
    }
    out[0] = creal(total); out[1] = cimag(total);
}

/* src/local_densitymatrix.hpp:134-164 ; targs/pairTargs in matching order, k ascending per output */
static void local_partialTrace(const orc_state* in, orc_state* out, const int* targs, const int* pairTargs, int nt) {
    int all[128]; memcpy(all, targs, sizeof(int) * nt); memcpy(all + nt, pairTargs, sizeof(int) * nt);
    qsort(all, 2 * nt, sizeof(int), cmp_int);
    idx_t traced = pow2(nt);
    for (int r = 0; r < in->numNodes; r++) {
        #pragma omp parallel for
        for (idx_t l = 0; l < out->numAmpsPerNode; l++) {
            idx_t base = insert_bits(l, all, 2 * nt, 0);
            amp_t acc = 0.;
            for (idx_t k = 0; k < traced; k++) acc += in->amps[r][set_bits(set_bits(base, targs, nt, k), pairTargs, nt, k)];
            out->amps[r][l] = acc;
        }
    }
}

/* src/misc.hpp:84-103 */
static int next_left_zero(idx_t mask, int b) { b--; while (bit_of(mask, b)) b--; return b; }
void orc_plan_partialTrace_targets(const int* ext, int n, int suffixSize, int* reordered) {
    idx_t mask = mask_of(ext, n);
    int maxFree = next_left_zero(mask, suffixSize);
    for (int q = n; q-- != 0; ) {
        if (ext[q] < suffixSize) reordered[q] = ext[q];
        else { reordered[q] = maxFree; maxFree = next_left_zero(mask, maxFree); }
    }
}
/* src/misc.hpp:106-135 */
void orc_plan_partialTrace_remaining(int numAll, const int* orig, const int* reordered, int n, int* remaining) {
    int all[128];
    for (int q = 0; q < numAll; q++) all[q] = q;
    for (int q = 0; q < n; q++) if (orig[q] != reordered[q]) { int t = all[orig[q]]; all[orig[q]] = all[reordered[q]]; all[reordered[q]] = t; }
    idx_t reMask = mask_of(reordered, n);
    int cnt = 0;
    for (int i = 0; i < numAll; i++) if (!bit_of(reMask, i)) remaining[cnt++] = all[i];
    idx_t remMask = mask_of(remaining, cnt);
    for (int i = 0; i < cnt; i++) { int p = remaining[i], q = p; for (int b = 0; b < p; b++) q -= !bit_of(remMask, b); remaining[i] = q; }
}

/* src/distributed_densitymatrix.hpp:347-407 ; returns a new state, MUTATES `in` on the relocation path */
orc_state* orc_dm_partialTrace(orc_state* in, const int* targetsIn, int nt) {
    int N = in->numQubits, L = in->logNumAmpsPerNode;
    if (N - nt < in->logNumNodes) return NULL;                         /* :352 */
    int targets[64]; sorted_copy(targets, targetsIn, nt);
    orc_state* out = orc_create(1, N - nt, in->numNodes);
    int pairs[64]; shifted(pairs, targets, nt, N);
    if (targets[nt - 1] + N < L) { local_partialTrace(in, out, targets, pairs, nt); return out; }

    int ext[128]; memcpy(ext, targets, sizeof(int) * nt); memcpy(ext + nt, pairs, sizeof(int) * nt);
    int re[128]; orc_plan_partialTrace_targets(ext, 2 * nt, L, re);
    for (int q = 2 * nt; q-- != 0; ) if (re[q] != ext[q]) orc_sv_swapGate(in, re[q], ext[q]);
    local_partialTrace(in, out, targets, re + nt, nt);
    int remaining[128]; orc_plan_partialTrace_remaining(2 * N, ext, re, 2 * nt, remaining);
    int cnt = 2 * N - 2 * nt;
    for (int q = cnt; q-- != 0; ) {
        if (remaining[q] == q) continue;
        int p = 0; while (remaining[p] != q) p++;
        orc_sv_swapGate(out, q, p);
        int t = remaining[q]; remaining[q] = remaining[p]; remaining[p] = t;
    }
    return out;
}
