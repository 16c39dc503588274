/*
 * dfsa_b200.h -- C-ABI of libdfsa_b200.so: the B200 (sm_100a) device layer behind the reference's
 * distributed full-state API.  Plain pointers and sizes only; no C++/torch types.
 *
 * The reference (TysonRayJones/Distributed-Full-State-Algorithms) is header-only C++; its process boundary
 * is MPI (src/communication.hpp) and its hot loops are OpenMP (src/local_*.hpp and the inline loops of
 * src/distributed_*.hpp).  This header is what a host program binds INSTEAD of those loops and of MPI:
 *   - dfsa_comm_*   replaces src/communication.hpp:16-47,172-177 (init/rank/size/barrier/reduce)
 *   - dfsa_state_*  replaces the std::vector storage of src/states.hpp:13-69 (amps + equal-size buffer, now in HBM)
 *   - dfsa_x_*      replaces comm_exchangeArrays / comm_asynchSendArray / comm_receiveArray
 *                   (src/communication.hpp:77-164): pairwise amplitude exchange, partner = rank XOR mask
 *   - dfsa_k_*      one entry per OpenMP loop of the reference, run as CUDA kernels
 * The host-side dispatch (local vs. exchange, relocation planning) stays C++ and lives in the drop-in
 * headers under distributed-full-state-algorithms_b200/host/, which call only the functions declared here.
 *
 * Conventions: amplitudes are interleaved (re,im) doubles, 16 B each (Amp = std::complex<double>,
 * src/types.hpp:26,37); `idx`/`num` arguments are in amplitudes; qubit q = bit q of the global amplitude
 * index, rank = top log2(P) bits (src/states.hpp:41-44).  Array selectors: DFSA_AMPS / DFSA_BUFFER.
 * Every function returns 0 on success or a negative dfsa_status; dfsa_last_error() gives the text.
 * All kernels are enqueued on the library's compute stream and are asynchronous; dfsa_comm_barrier(),
 * dfsa_device_sync(), downloads and reductions synchronise.  There is NO CPU fallback: without a
 * CUDA device every call fails with DFSA_ERR_CUDA.
 */
#ifndef DFSA_B200_H
#define DFSA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dfsa_state dfsa_state;

enum dfsa_status {
    DFSA_OK = 0,
    DFSA_ERR_CUDA = -1,        /* a CUDA runtime call failed (or no device) */
    DFSA_ERR_NCCL = -2,
    DFSA_ERR_ARG = -3,         /* precondition violated (the reference would assert) */
    DFSA_ERR_COMM = -4,        /* bootstrap / transport failure */
    DFSA_ERR_UNSUPPORTED = -5
};

enum { DFSA_AMPS = 0, DFSA_BUFFER = 1 };
enum { DFSA_DEPOL2_CORRECTED = 16 };  /* OR-ed into the `phase` of dfsa_k_depol2Pair / dfsa_k_depol2Quad: the true channel, not the reference's formulas */
enum { DFSA_MAX_QUBITS = 64 };

const char* dfsa_last_error(void);
const char* dfsa_version(void);

/* ---- communication environment: src/communication.hpp:16-47 ------------------------------------------- */
/* Bootstrap from the environment. RANK/WORLD_SIZE(/LOCAL_RANK) set (torchrun-style launch): join that job.
 * DFSA_NP=P set: fork P-1 children now (must be the first CUDA-touching call of the process, like MPI_Init
 * being the first statement of the reference's mains). Neither: single rank.  Idempotent. */
int dfsa_comm_init(void);
/* Bootstrap with an externally distributed NCCL unique id (128 bytes from dfsa_comm_get_unique_id on rank 0),
 * e.g. broadcast by torch.distributed. `device` < 0 means rank % deviceCount. */
int dfsa_comm_get_unique_id(void* out128);
int dfsa_comm_init_with_id(int rank, int numRanks, const void* uniqueId128, int device);
int dfsa_comm_finalize(void);                      /* comm_end(): barrier, then tear down */
int dfsa_comm_rank(void);                          /* comm_getRank() */
int dfsa_comm_size(void);                          /* comm_getNumNodes() */
int dfsa_comm_barrier(void);                       /* comm_synch(): device sync + inter-rank barrier */
int dfsa_device_sync(void);
const char* dfsa_comm_transport(void);             /* "single", "nccl" or "ipc" */
/* Collective. 1: exchanges fused into the kernels that consume them (peer shards read over NVLink), 0: the staged pack /
 * exchange / combine path of the reference, -1: what DFSA_FUSED_EXCHANGE says (default fused where every rank can map every
 * other rank's shards). dfsa_comm_fused_active(): 0 staged, 1 fused with host synchronisation, 2 fused and stream-ordered. */
int dfsa_comm_set_fused(int mode);
int dfsa_comm_fused_active(void);
/* Measurement: device time of the kernel of this rank's most recent fused exchange step (stream-ordered mode), taken after the
 * partners' READY signals -- i.e. without the time spent waiting for a rank that arrives late. -1 if none. Synchronises. */
int dfsa_comm_last_exchange_ms(double* ms);
void* dfsa_stream_compute(void);                   /* cudaStream_t the kernels run on (for CUDA-event timing) */

/* ---- measurement helpers (no counterpart in the reference, which times with std::chrono around comm_synch, main.cpp:28-35) */
/* CUDA events recorded on the compute stream, so a timed region brackets exactly the kernels enqueued between them */
int dfsa_event_create(void** event);
int dfsa_event_record(void* event);
int dfsa_event_elapsed_ms(void* start, void* stop, double* ms);   /* synchronises on `stop` */
int dfsa_event_destroy(void* event);
uint64_t dfsa_launch_count(void);                  /* kernels launched by this library so far (this process) */
/* page-locked host memory for full-bandwidth state upload / download */
int dfsa_host_alloc_pinned(uint64_t bytes, void** out);
int dfsa_host_free_pinned(void* ptr);

/* ---- state storage: src/states.hpp:13-69 --------------------------------------------------------------- */
/* Collective. Allocates this rank's shard (2^(n-k) or 2^(2N-k) amps) and, when P>1, the equal-size
 * exchange buffer, both zero-filled, in HBM. Fails with DFSA_ERR_ARG if 2^numQubits < P (states.hpp:35). */
int dfsa_state_create(int isDensity, unsigned numQubits, dfsa_state** out);
int dfsa_state_destroy(dfsa_state* s);
double*  dfsa_state_ptr(dfsa_state* s, int which);                /* device pointer of DFSA_AMPS / DFSA_BUFFER */
uint64_t dfsa_state_num_amps_per_node(const dfsa_state* s);
unsigned dfsa_state_log_num_amps_per_node(const dfsa_state* s);
unsigned dfsa_state_num_qubits(const dfsa_state* s);
int      dfsa_state_is_density(const dfsa_state* s);
int dfsa_state_swap_arrays(dfsa_state* s);                        /* amps <-> buffer pointer swap */
/* host <-> device copies of this rank's shard, [first, first+num) in LOCAL amplitude indices */
int dfsa_state_upload(dfsa_state* s, int which, uint64_t first, uint64_t num, const double* host);
int dfsa_state_download(dfsa_state* s, int which, uint64_t first, uint64_t num, double* host);
/* whole state (all ranks, global order) into host memory of every rank: getAllVecAmps (test_utilities.hpp:419) */
int dfsa_state_download_all(dfsa_state* s, double* hostAll);
int dfsa_state_upload_all(dfsa_state* s, const double* hostAll);  /* each rank keeps its slice */
int dfsa_state_init_zero(dfsa_state* s);
int dfsa_state_init_hash(dfsa_state* s, uint64_t seed);           /* synthetic state, SURVEY 8(d) */
int dfsa_state_norm2(dfsa_state* s, double* out);                 /* sum |amp|^2 over all ranks */
/* Device-resident test utilities (the reference's are host loops over gathered copies,
 * tests/test_utilities.hpp:419-524). All collective, results identical on every rank. */
int dfsa_state_init_plus(dfsa_state* s);                          /* every amplitude = 2^(-n/2) (sv) or 2^(-N) (dm: |+><+|) */
int dfsa_state_copy(dfsa_state* dst, const dfsa_state* src);      /* dst.amps = src.amps (same shape) */
/* agreesWith (test_utilities.hpp:470-487) without leaving the device, two-sided: *maxAbsDiff = max over all amplitudes and
 * ranks of max(|re a - re b|, |im a - im b|) (NaN anywhere -> NaN), *numUnequal = how many amplitudes differ in value
 * (== on both doubles, so +0 == -0: the bit-exact criterion of SURVEY 8c), *maxAbsRef = max |component| of b. */
int dfsa_state_compare(dfsa_state* a, dfsa_state* b, double* maxAbsDiff, uint64_t* numUnequal, double* maxAbsRef);
/* the same against the synthetic state dfsa_state_init_hash(seed) would produce, regenerated on the fly (no second shard) */
int dfsa_state_compare_hash(dfsa_state* s, uint64_t seed, double* maxAbsDiff, uint64_t* numUnequal, double* maxAbsRef);

/* ---- pairwise exchange: src/communication.hpp:77-164 ---------------------------------------------------- */
/* comm_exchangeArrays(toSend, sendStart, toReceive, recvStart, num, pairRank): both partners call it with the
 * same recv array/offset; data lands in the partner's `recvWhich` array at `recvStart`. Ordered after all
 * previously enqueued kernels; later kernels are ordered after the received data. */
int dfsa_x_exchange(dfsa_state* s, int sendWhich, uint64_t sendStart, int recvWhich, uint64_t recvStart,
                    uint64_t num, int pairRank);
/* one-directional pair used by damping (src/distributed_densitymatrix.hpp:292,306): sender / receiver side */
int dfsa_x_send(dfsa_state* s, int sendWhich, uint64_t sendStart, int recvWhich, uint64_t recvStart, uint64_t num, int pairRank);
int dfsa_x_recv(dfsa_state* s, int recvWhich, uint64_t recvStart, uint64_t num, int pairRank);
int dfsa_x_allreduce_amp(double reim[2]);                         /* comm_reduceAmp, host value in/out */
/* Fused + pipelined forms of "exchange the whole shard, then combine" (distributed_statevector.hpp:26-38 and :227-241):
 * the shard travels in chunks on the comm stream while the combine kernel of the previous chunk runs on the compute
 * stream, so a prefix gate costs ~max(NVLink time, HBM time) instead of their sum. Same results as
 * dfsa_x_exchange + dfsa_k_combine / dfsa_k_pauliCombine. */
int dfsa_xk_exchangeCombine(dfsa_state* s, int pairRank, const double f0[2], const double f1[2]);
/* swapGate of a suffix qubit `qb1` with the prefix qubit whose rank bit distinguishes this rank from `pairRank`
 * (distributed_statevector.hpp:140-186). `movingBit` is the value of qb1 in the half that leaves this rank (= NOT this
 * rank's bit of the prefix qubit). With peer-mapped shards: ONE out-of-place pass -- buffer[j] = amps[j] where bit qb1 of
 * j stays, else the partner's amps[j ^ (1 << qb1)] read over NVLink -- then amps <-> buffer. Otherwise the reference's
 * steps: contiguous half exchange + copy (qb1 top suffix qubit) or pack + dfsa_x_exchange + unpack. */
int dfsa_xk_swapSuffixPrefix(dfsa_state* s, unsigned qb1, unsigned movingBit, int pairRank);
/* Relocation of manyTargGate's prefix targets (distributed_statevector.hpp:193-223): swap suffix qubit suffixQubits[i] with
 * prefix qubit prefixQubits[i] for every i (numPairs <= 4). COLLECTIVE: every rank calls it. One pair = the fused swap above;
 * several pairs with peer-mapped shards = one gather pass over the 2^numPairs shards of the rank's group; otherwise the
 * reference's sequence of swaps. Its own inverse. */
int dfsa_xk_relocate(dfsa_state* s, const uint32_t* suffixQubits, const uint32_t* prefixQubits, unsigned numPairs);
/* Host-only (no device needed): who supplies what in the single-shot relocation -- owners[sigma] = the rank whose shard holds
 * the amplitudes that land at suffix bits sigma (bit i of sigma <-> pair i) on `rank`, -1 beyond 2^numPairs; *rho = rank's
 * own bits of the swapped prefix qubits. new[j] = shard(owners[sigma(j)])[j with its landing bits := rho]. */
int dfsa_plan_relocate(int rank, unsigned logNumAmps, const uint32_t* prefixQubits, unsigned numPairs, int owners[16], unsigned* rho);
/* oneQubitDepolarising / damping on a qubit whose bra bit is a rank bit (distributed_densitymatrix.hpp:110-141, :284-317):
 * pack + half exchange (one-way for damping) + combine of the reference. `bit` = this rank's bit of that qubit. With
 * peer-mapped shards: one out-of-place pass that reads the partner's half over NVLink, then amps <-> buffer. */
int dfsa_xk_depol1Prefix(dfsa_state* s, unsigned qb, unsigned bit, double prob, int pairRank);
int dfsa_xk_dampingPrefix(dfsa_state* s, unsigned qb, unsigned bit, double prob, int pairRank);
/* manyCtrlOneTargGate with a prefix target and suffix controls (distributed_statevector.hpp:43-78): amps[k] = f0*amps[k] +
 * f1*partner_amps[k] on the sub-cube where every control in suffixCtrls (strictly increasing) is 1. */
int dfsa_xk_ctrlPrefixTarg(dfsa_state* s, const uint32_t* suffixCtrls, unsigned numCtrls, int pairRank, const double f0[2], const double f1[2]);
/* twoQubitDepolarising with one (pair, :146-183) or both (quad, :187-237) bra bits in the rank index; qb1 < qb2; bit / bit0 /
 * bit1 = this rank's bits of those qubits. corrected = 0: the reference's formulas, literally. */
int dfsa_xk_depol2Pair(dfsa_state* s, unsigned qb1, unsigned qb2, unsigned bit, double prob, int corrected, int pairRank);
int dfsa_xk_depol2Quad(dfsa_state* s, unsigned qb1, unsigned qb2, unsigned bit0, unsigned bit1, double prob, int corrected, int pairRank0, int pairRank1);
/* Measurement only: this rank pulls pairRank's whole shard into its exchange buffer (all pairs at once, both directions);
 * mode 0 = remote loads from a kernel, 1 = copy engine, 2 = remote loads, even ranks only (one-way traffic). *ms = device time.
 * NVLink GB/s per direction = 16 * A / ms. */
int dfsa_xk_measure_link(dfsa_state* s, int pairRank, int mode, double* ms);
int dfsa_xk_exchangePauliCombine(dfsa_state* s, int pairRank, uint64_t maskXY, uint64_t maskYZ, unsigned numY,
                                 const double f[2], const double g[2], int exact);

/* ---- state-vector kernels (src/local_statevector.hpp, inline loops of src/distributed_statevector.hpp) -- */
/* gate pointers are HOST pointers to row-major interleaved complex doubles */
/* K1+K2: 2x2 gate on `target` where all `ctrls` are 1 (numCtrls may be 0). local_statevector.hpp:14,32 */
int dfsa_k_ctrlOneTarg(dfsa_state* s, const uint32_t* ctrls, unsigned numCtrls, unsigned target, const double gate[8]);
/* K1/K2 for a RUN of gates: gates[i] is a 2x2 gate (row-major interleaved complex) on suffix bit `target`, applied where every
 * bit of ctrlMask (a mask on the GLOBAL index: suffix and rank bits) is 1. Applied in order; bit-identical to numGates calls of
 * dfsa_k_ctrlOneTarg, but consecutive gates whose targets fit one shared-memory tile (bits 0..3 plus up to seven more) cross HBM
 * once together. DFSA_FUSE_GATES=0: one pass per gate. dfsa_plan_gateSequence: the batching, host-only (numBatches <= numGates;
 * tileBitsOut holds 11 bits per batch, groupBitsOut 3 per gate). */
typedef struct dfsa_gate1 { double matrix[8]; uint64_t ctrlMask; uint32_t target; uint32_t reserved; } dfsa_gate1;
int dfsa_k_gateSequence(dfsa_state* s, const dfsa_gate1* gates, unsigned numGates);
int dfsa_plan_gateSequence(const dfsa_gate1* gates, unsigned numGates, unsigned logNumAmps, uint32_t* batchOfGate, uint32_t* groupOfGate,
                           uint32_t* batchIsTiled, uint32_t* tileBitsOut, uint32_t* groupBitsOut, unsigned* numBatches);
/* K3: swap of two suffix qubits. local_statevector.hpp:54 */
int dfsa_k_swap(dfsa_state* s, unsigned qb1, unsigned qb2);
/* K4: dense 2^t x 2^t gate on suffix targets, gate bit i <-> targets[i]. local_statevector.hpp:72 */
int dfsa_k_manyTarg(dfsa_state* s, const uint32_t* targets, unsigned numTargets, const double* gate);
/* krausMap after relocation (src/distributed_densitymatrix.hpp:79-89 + getSuperoperator, src/misc.hpp:58-81): applies
 * sum_K conj(K) (x) K as a gate on the 2t suffix bits targets2t = {targets, targets + N} (gate bit i <-> targets2t[i]).
 * krausOps: numOps host matrices 2^t x 2^t, row-major interleaved. From t = 4 the superoperator is built on the device. */
int dfsa_k_krausMap(dfsa_state* s, const uint32_t* targets2t, unsigned numTargets2t, const double* krausOps, unsigned numOps);
/* Host-only (no device needed): the tile plan of the tensor-core manyTarg kernel (3 <= numTargets <= 6, logNumAmps >= 9) --
 * the 9 index bits of a tile (ascending), their roles (< numTargets: gate-row bit i = targets[i], else vector bit
 * role - numTargets), the slab byte-offset contribution of each tile bit (address order, XOR-swizzled) and the same per
 * gate-row bit / vector bit. tests/test_manytarg_layout.py checks on CPU that every fragment access pattern is
 * shared-memory bank-conflict free for every target placement. */
int dfsa_plan_manyTargLayout(const uint32_t* targets, unsigned numTargets, unsigned logNumAmps, uint32_t tileBits[9],
                             uint32_t roles[9], uint32_t bitOff[9], uint32_t rowBit[6], uint32_t colBit[6]);
/* K5: amps[j0] = f*amps[j0] + g*b(j1)*amps[j1], j1 = j0^maskXY, b = i^numY * (-1)^parity(global(j1) & maskYZ);
 * maskXY==0 is the diagonal case. `exact` selects the move/negate-only path (f=0,g=1: pauliTensor).
 * local_statevector.hpp:102 */
int dfsa_k_pauli(dfsa_state* s, uint64_t maskXY, uint64_t maskYZ, unsigned numY, const double f[2], const double g[2], int exact);
/* K6: amps[j] *= exp(+-i theta) by parity of (global index & targMask). local_statevector.hpp:138 */
int dfsa_k_phase(dfsa_state* s, uint64_t targMask, double theta);
/* K7: amps[i] = f0*amps[i] + f1*buffer[i]. distributed_statevector.hpp:36-38 */
int dfsa_k_combine(dfsa_state* s, const double f0[2], const double f1[2]);
/* K8/K10/K19-K22 building blocks on sub-cubes: local index k = insert `values` bits at sorted `positions` into j.
 *   pack:    buffer[dstStart + j] = amps[k]                          (distributed_statevector.hpp:56,169)
 *   unpack:  amps[k] = buffer[srcStart + j]                          (distributed_statevector.hpp:180)
 *   combine: amps[k] = f0*amps[k] + f1*buffer[srcStart + j]          (distributed_statevector.hpp:72) */
int dfsa_k_pack(dfsa_state* s, const uint32_t* positions, unsigned numPositions, uint64_t values, uint64_t dstStart);
int dfsa_k_unpack(dfsa_state* s, const uint32_t* positions, unsigned numPositions, uint64_t values, uint64_t srcStart);
int dfsa_k_combineSub(dfsa_state* s, const uint32_t* positions, unsigned numPositions, uint64_t values, uint64_t srcStart,
                      const double f0[2], const double f1[2]);
/* K9: amps[dstStart..+num) = buffer[srcStart..+num). distributed_statevector.hpp:133,152 */
int dfsa_k_copyFromBuffer(dfsa_state* s, uint64_t dstStart, uint64_t srcStart, uint64_t num);
/* K11: amps[j0] = f*amps[j0] + g*b*buffer[j0^maskXY], sign from pairRank's global index. distributed_statevector.hpp:227 */
int dfsa_k_pauliCombine(dfsa_state* s, int pairRank, uint64_t maskXY, uint64_t maskYZ, unsigned numY, const double f[2], const double g[2], int exact);
/* K18: amps *= factor (complex). distributed_densitymatrix.hpp:46 */
int dfsa_k_scaleAll(dfsa_state* s, const double factor[2]);

/* ---- density-matrix kernels (src/local_densitymatrix.hpp, inline loops of src/distributed_densitymatrix.hpp) */
int dfsa_k_oneQubitDephasing(dfsa_state* s, unsigned qb, double prob);                 /* K12, local_densitymatrix.hpp:12 */
int dfsa_k_twoQubitDephasing(dfsa_state* s, unsigned qb1, unsigned qb2, double prob);  /* K13, :45 */
int dfsa_k_oneQubitDepolarising(dfsa_state* s, unsigned qb, double prob);              /* K14, :63 (suffix case) */
int dfsa_k_twoQubitDepolarising(dfsa_state* s, unsigned qb1, unsigned qb2, double prob, int corrected); /* K15, :83 */
int dfsa_k_damping(dfsa_state* s, unsigned qb, double prob);                           /* K16, :111 (suffix case) */
/* K17: out.amps[l] = sum_k in.amps[...]; targets/pairTargets are suffix bit positions in matching order. :134 */
int dfsa_k_partialTrace(dfsa_state* in, dfsa_state* out, const uint32_t* targets, const uint32_t* pairTargets, unsigned numTargets);
/* K19: after the half exchange: scale the non-exchanged half by c3, amps[k] = c2*amps[k] + c1*buffer[A/2 + j].
 * distributed_densitymatrix.hpp:130-141 */
int dfsa_k_depol1Combine(dfsa_state* s, unsigned qb, unsigned bit, double prob);
/* K20/K21: the three phases of the prefix twoQubitDepolarising branches, formulas as in the reference.
 * distributed_densitymatrix.hpp:152-183 (pair), :195-237 (quad); phase = 0,1,2 */
int dfsa_k_depol2Pair(dfsa_state* s, unsigned q0, unsigned q1, unsigned q2, unsigned bit, double prob, int phase);
int dfsa_k_depol2Quad(dfsa_state* s, unsigned q0, unsigned q1, unsigned bit0, unsigned bit1, double prob, int phase);
/* K22: damping across ranks. phase 0 (bit=1 ranks): buffer[j] = amps[k], amps[k] *= 1-p; phase 1 (all): other half *= sqrt(1-p);
 * phase 2 (bit=0 ranks): amps[k] += p*buffer[j]. distributed_densitymatrix.hpp:284-313 */
int dfsa_k_dampingPrefix(dfsa_state* s, unsigned qb, unsigned bit, double prob, int phase);
/* K23: local part of sum_t coeff_t Tr(P_t rho); paulis is numTerms x N (codes 0..3, qubit q of term t at [t*N+q]).
 * Result (this rank's partial sum) is written to out[2]; combine with dfsa_x_allreduce_amp. :322 */
int dfsa_k_expecPauliString(dfsa_state* s, const double* coeffs, unsigned numTerms, const uint32_t* paulis, double out[2]);
/* K23 + X11 (comm_reduceAmp, src/communication.hpp:172) in one call: where the ranks share a node the reduction kernel of every
 * rank publishes into the job's shared page and each host sums all slots in rank order -- *outIsGlobal = 1, out[] is the value
 * of distributed_densitymatrix_expecPauliString on every rank. Otherwise *outIsGlobal = 0 and out[] is the local part. */
int dfsa_kx_expecPauliString(dfsa_state* s, const double* coeffs, unsigned numTerms, const uint32_t* paulis, double out[2], int* outIsGlobal);

#ifdef __cplusplus
}
#endif
#endif /* DFSA_B200_H */
