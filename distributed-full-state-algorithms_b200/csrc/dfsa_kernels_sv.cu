// dfsa_kernels_sv.cu -- state-vector kernels K1-K11, K18 for sm_100a.
// Each extern "C" entry cites the reference loop it replaces; see include/dfsa_b200.h for the contract.
#include <algorithm>
#include <vector>

#include "dfsa_stream_kernels.cuh"

static inline uint64_t rankShiftOf(const dfsa_state* s, int rank) { return (uint64_t)rank << s->logNumAmps; }

// ---------------------------------------------------------------------------------------------------------
// K1 + K2: local_statevector.hpp:14-29 (oneTarg) and :32-51 (manyCtrlOneTarg).
// item j -> i1 = insert ones at sorted(ctrls U {target}); i0 = i1 with the target bit cleared.
// HBM traffic: 32 B per touched amplitude pair half, i.e. 32*A/2^c bytes for c controls.
template <int NPOS>
static int launchCtrlOneTarg(double2* amps, uint64_t items, const BitSpec& spec, uint64_t ones, uint64_t targBit, const Gate2& g) {
    using Item = PairAt;   // idx = i1
    auto ld = [=] __device__(uint64_t j) {
        Item it;
        uint64_t v = (NPOS > 0) ? insertZeroBitsN<(NPOS > 0 ? NPOS : 1)>(j, spec) : insertZeroBits(j, spec);
        it.idx = v | ones;
        it.a1 = amps[it.idx];
        it.a0 = amps[it.idx ^ targBit];
        return it;
    };
    auto st = [=] __device__(uint64_t, const Item& it) {
        amps[it.idx ^ targBit] = cfma(g.m01, it.a1, cmul(g.m00, it.a0));
        amps[it.idx]           = cfma(g.m11, it.a1, cmul(g.m10, it.a0));
    };
    return launchStream<1, Item>(items, ld, st);
}

extern "C" int dfsa_k_ctrlOneTarg(dfsa_state* s, const uint32_t* ctrls, unsigned numCtrls, unsigned target, const double gate[8]) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s && gate, "null argument");
    DFSA_REQUIRE(numCtrls + 1 <= s->logNumAmps, "more qubits than local bits");
    std::vector<uint32_t> qs(ctrls, ctrls + numCtrls);
    qs.push_back(target);
    BitSpec spec;
    uint64_t ones;
    DFSA_TRY(sortedSpec(qs.data(), numCtrls + 1, s->logNumAmps, &spec, &ones));
    Gate2 g{hostAmp(gate), hostAmp(gate + 2), hostAmp(gate + 4), hostAmp(gate + 6)};
    uint64_t items = s->numAmps >> (numCtrls + 1);
    uint64_t targBit = 1ULL << target;
    double2* a = s->arr[DFSA_AMPS];
    switch (spec.n) {
        case 1:  return launchCtrlOneTarg<1>(a, items, spec, ones, targBit, g);
        case 2:  return launchCtrlOneTarg<2>(a, items, spec, ones, targBit, g);
        case 3:  return launchCtrlOneTarg<3>(a, items, spec, ones, targBit, g);
        case 4:  return launchCtrlOneTarg<4>(a, items, spec, ones, targBit, g);
        default: return launchCtrlOneTarg<0>(a, items, spec, ones, targBit, g);
    }
}

// ---------------------------------------------------------------------------------------------------------
// K3: local_statevector.hpp:54-69. Pure permutation (bit-exact): amps[..01..] <-> amps[..10..]. 16*A bytes.
extern "C" int dfsa_k_swap(dfsa_state* s, unsigned qb1, unsigned qb2) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s, "null state");
    DFSA_REQUIRE(qb1 != qb2 && qb1 < s->logNumAmps && qb2 < s->logNumAmps, "swap needs two distinct suffix qubits");
    unsigned lo = std::min(qb1, qb2), hi = std::max(qb1, qb2);
    uint64_t bLo = 1ULL << lo, bHi = 1ULL << hi;
    double2* amps = s->arr[DFSA_AMPS];
    using Item = MovePair;   // a0 = amp at ..01.., a1 = amp at ..10.., idx = j10
    auto ld = [=] __device__(uint64_t k) {
        Item it;
        it.idx = insertZeroBit(insertZeroBit(k, lo), hi) | bHi;       // hi qubit = 1, lo qubit = 0
        it.a1 = amps[it.idx];
        it.a0 = amps[it.idx ^ bHi ^ bLo];
        return it;
    };
    auto st = [=] __device__(uint64_t, const Item& it) {
        amps[it.idx] = it.a0;
        amps[it.idx ^ bHi ^ bLo] = it.a1;
    };
    return launchStream<2, Item>(s->numAmps >> 2, ld, st);
}

// ---------------------------------------------------------------------------------------------------------
// diagonal-by-parity kernel: amps[j] *= (parity(global(j) & mask) ? c1 : c0).
// K6 (local_statevector.hpp:138-153) and the X/Y-free Pauli case.
static int launchDiagParity(dfsa_state* s, uint64_t mask, double2 c0, double2 c1) {
    double2* amps = s->arr[DFSA_AMPS];
    uint64_t rs = rankShiftOf(s, s->rank);
    auto ld = [=] __device__(uint64_t j) { return Amp1{amps[j]}; };
    auto st = [=] __device__(uint64_t j, const Amp1& v) { amps[j] = cmul(v.a, parity64((rs | j) & mask) ? c1 : c0); };
    return launchStream<2, Amp1>(s->numAmps, ld, st);
}

extern "C" int dfsa_k_phase(dfsa_state* s, uint64_t targMask, double theta) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s, "null state");
    double c = cos(theta), sn = sin(theta);
    return launchDiagParity(s, targMask, make_double2(c, sn), make_double2(c, -sn));
}

// ---------------------------------------------------------------------------------------------------------
// K5: local_statevector.hpp:102-135. Pairs (j0, j1 = j0 ^ maskXY), j0 has the highest X/Y bit clear.
//   amps[j0] = f*a0 + (g*b1)*a1,  amps[j1] = f*a1 + (g*b0)*a0,  b = i^numY * (-1)^parity(global & maskYZ)
// EXACT (pauliTensor, f=0 g=1): pure move + multiplication by a power of i -> bit-exact.
template <bool EXACT>
static int launchPauliPairs(dfsa_state* s, uint64_t maskXY, uint64_t maskYZ, unsigned numY, double2 f, double2 h) {
    double2* amps = s->arr[DFSA_AMPS];
    uint64_t rs = rankShiftOf(s, s->rank);
    unsigned hiPos = 63u - (unsigned)__builtin_clzll(maskXY);
    using Item = PairAt;   // idx = j0
    auto ld = [=] __device__(uint64_t j) {
        Item it;
        it.idx = insertZeroBit(j, hiPos);
        it.a0 = amps[it.idx];
        it.a1 = amps[it.idx ^ maskXY];
        return it;
    };
    auto st = [=] __device__(uint64_t, const Item& it) {
        uint64_t j1 = it.idx ^ maskXY;
        unsigned p0 = parity64((rs | it.idx) & maskYZ), p1 = parity64((rs | j1) & maskYZ);
        if (EXACT) {
            amps[it.idx] = mulPowI(it.a1, numY + 2u * p1);
            amps[j1]    = mulPowI(it.a0, numY + 2u * p0);
        } else {
            double2 h1 = p1 ? make_double2(-h.x, -h.y) : h, h0 = p0 ? make_double2(-h.x, -h.y) : h;
            amps[it.idx] = cfma(h1, it.a1, cmul(f, it.a0));
            amps[j1]    = cfma(h0, it.a0, cmul(f, it.a1));
        }
    };
    return launchStream<1, Item>(s->numAmps >> 1, ld, st);
}

double2 dfsaPowIHost(unsigned k);
static double2 powIHost(unsigned k) {
    switch (k & 3u) { case 0: return make_double2(1, 0); case 1: return make_double2(0, 1); case 2: return make_double2(-1, 0); default: return make_double2(0, -1); }
}
static double2 cmulHost(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

extern "C" int dfsa_k_pauli(dfsa_state* s, uint64_t maskXY, uint64_t maskYZ, unsigned numY, const double f[2], const double g[2], int exact) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s && f && g, "null argument");
    DFSA_REQUIRE((maskXY >> s->logNumAmps) == 0, "maskXY must only hold suffix bits");
    double2 ff = hostAmp(f), h = cmulHost(hostAmp(g), powIHost(numY));
    if (maskXY == 0) {
        // No X/Y on this shard: the operator is diagonal, amps[j] *= f + (+-)g*i^numY.  (The reference's local
        // loop has 2^0/2 = 0 inner iterations here and silently does nothing; this build applies the operator.)
        if (exact) {
            double2 p = powIHost(numY);
            return launchDiagParity(s, maskYZ, p, make_double2(-p.x, -p.y));
        }
        return launchDiagParity(s, maskYZ, make_double2(ff.x + h.x, ff.y + h.y), make_double2(ff.x - h.x, ff.y - h.y));
    }
    return exact ? launchPauliPairs<true>(s, maskXY, maskYZ, numY, ff, h) : launchPauliPairs<false>(s, maskXY, maskYZ, numY, ff, h);
}

// K11: distributed_statevector.hpp:227-241. amps[j0] = f*amps[j0] + g*b1*buffer[j0 ^ maskXY], sign from the
// partner's global index. 48*A bytes (read amps, read buffer, write amps). Range form: j0 in [first, first+num).
template <bool EXACT>
static int launchPauliCombine(dfsa_state* s, uint64_t first, uint64_t num, int pairRank, uint64_t maskXY, uint64_t maskYZ, unsigned numY, double2 f, double2 h) {
    double2* amps = s->arr[DFSA_AMPS];
    const double2* buf = s->arr[DFSA_BUFFER];
    uint64_t rs = rankShiftOf(s, pairRank);
    auto ld = [=] __device__(uint64_t k) { uint64_t j0 = first + k; return Amp2{amps[j0], buf[j0 ^ maskXY]}; };
    auto st = [=] __device__(uint64_t k, const Amp2& v) {
        uint64_t j0 = first + k;
        unsigned p1 = parity64((rs | (j0 ^ maskXY)) & maskYZ);
        if (EXACT) amps[j0] = mulPowI(v.a1, numY + 2u * p1);
        else {
            double2 h1 = p1 ? make_double2(-h.x, -h.y) : h;
            amps[j0] = cfma(h1, v.a1, cmul(f, v.a0));
        }
    };
    return launchStream<1, Amp2>(num, ld, st);
}

int dfsaLaunchPauliCombineRange(dfsa_state* s, uint64_t first, uint64_t num, int pairRank, uint64_t maskXY, uint64_t maskYZ,
                                unsigned numY, double2 f, double2 h, bool exact) {
    return exact ? launchPauliCombine<true>(s, first, num, pairRank, maskXY, maskYZ, numY, f, h)
                 : launchPauliCombine<false>(s, first, num, pairRank, maskXY, maskYZ, numY, f, h);
}

extern "C" int dfsa_k_pauliCombine(dfsa_state* s, int pairRank, uint64_t maskXY, uint64_t maskYZ, unsigned numY, const double f[2], const double g[2], int exact) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s && f && g && s->arr[DFSA_BUFFER], "null argument / no exchange buffer");
    double2 ff = hostAmp(f), h = cmulHost(hostAmp(g), powIHost(numY));
    return dfsaLaunchPauliCombineRange(s, 0, s->numAmps, pairRank, maskXY, maskYZ, numY, ff, h, exact != 0);
}

// ---------------------------------------------------------------------------------------------------------
// K7: distributed_statevector.hpp:36-38. amps[i] = f0*amps[i] + f1*buffer[i]. 48*A bytes. Range form: i in [first, first+num).
// partnerFirst: the rank holding bit 1 of the target forms m10*partner first and then adds m11*own -- the FMA nesting the local
// kernel uses for its upper output (K1: cfma(m11, a1, cmul(m10, a0))), so a gate gives the same bits whether its qubit sits on
// a suffix bit or a rank bit (the lazy layout may move it between the two).
int dfsaLaunchCombineRange(dfsa_state* s, uint64_t first, uint64_t num, double2 c0, double2 c1, bool partnerFirst) {
    double2* amps = s->arr[DFSA_AMPS] + first;
    const double2* buf = s->arr[DFSA_BUFFER] + first;
    auto ld = [=] __device__(uint64_t i) { return Amp2{amps[i], buf[i]}; };
    auto st = [=] __device__(uint64_t i, const Amp2& v) { amps[i] = partnerFirst ? cfma(c0, v.a0, cmul(c1, v.a1)) : cfma(c1, v.a1, cmul(c0, v.a0)); };
    return launchStream<1, Amp2>(num, ld, st);
}

extern "C" int dfsa_k_combine(dfsa_state* s, const double f0[2], const double f1[2]) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s && f0 && f1 && s->arr[DFSA_BUFFER], "null argument / no exchange buffer");
    return dfsaLaunchCombineRange(s, 0, s->numAmps, hostAmp(f0), hostAmp(f1), false);
}

// ---------------------------------------------------------------------------------------------------------
// X1+K7 / X6+K11 as ONE kernel over peer memory: the partner's shard is mapped into this process
// (CUDA IPC; NVLink peer access between GPUs), so the combine loads the partner's amplitudes directly over NVLink while
// streaming its own shard from HBM, and writes the result out-of-place into `buffer`; the caller then swaps amps<->buffer.
// Per rank: 16*A bytes over NVLink, 32*A bytes of HBM traffic (the staged path moves 80*A through HBM), no staging pass.
int dfsaLaunchFusedCombine(dfsa_state* s, const double2* remote, double2 c0, double2 c1, bool partnerFirst) {
    const double2* amps = s->arr[DFSA_AMPS];
    double2* out = s->arr[DFSA_BUFFER];
    auto ld = [=] __device__(uint64_t i) { return Amp2{amps[i], remote[i]}; };
    auto st = [=] __device__(uint64_t i, const Amp2& v) { out[i] = partnerFirst ? cfma(c0, v.a0, cmul(c1, v.a1)) : cfma(c1, v.a1, cmul(c0, v.a0)); };
    return launchStreamRemote<Amp2>(s->numAmps, ld, st);
}

template <bool EXACT>
static int launchFusedPauli(dfsa_state* s, const double2* remote, int pairRank, uint64_t maskXY, uint64_t maskYZ, unsigned numY, double2 f, double2 h) {
    const double2* amps = s->arr[DFSA_AMPS];
    double2* out = s->arr[DFSA_BUFFER];
    uint64_t rs = rankShiftOf(s, pairRank);
    auto ld = [=] __device__(uint64_t j0) { return Amp2{amps[j0], remote[j0 ^ maskXY]}; };
    auto st = [=] __device__(uint64_t j0, const Amp2& v) {
        unsigned p1 = parity64((rs | (j0 ^ maskXY)) & maskYZ);
        if (EXACT) out[j0] = mulPowI(v.a1, numY + 2u * p1);
        else {
            double2 h1 = p1 ? make_double2(-h.x, -h.y) : h;
            out[j0] = cfma(h1, v.a1, cmul(f, v.a0));
        }
    };
    return launchStreamRemote<Amp2>(s->numAmps, ld, st);
}

int dfsaLaunchFusedPauliCombine(dfsa_state* s, const double2* remote, int pairRank, uint64_t maskXY, uint64_t maskYZ,
                                unsigned numY, double2 f, double2 h, bool exact) {
    return exact ? launchFusedPauli<true>(s, remote, pairRank, maskXY, maskYZ, numY, f, h)
                 : launchFusedPauli<false>(s, remote, pairRank, maskXY, maskYZ, numY, f, h);
}

// buffer = remote (the link microbenchmark of dfsa_xk_measure_link: a prefix gate's traffic without its arithmetic)
int dfsaLaunchPull(dfsa_state* s, const double2* remote) {
    double2* out = s->arr[DFSA_BUFFER];
    auto ld = [=] __device__(uint64_t i) { return Amp1{remote[i]}; };
    auto st = [=] __device__(uint64_t i, const Amp1& v) { out[i] = v.a; };
    return launchStreamRemote<Amp1>(s->numAmps, ld, st);
}

// K18: distributed_densitymatrix.hpp:44-49 (amp *= -1) generalised to a complex factor.
extern "C" int dfsa_k_scaleAll(dfsa_state* s, const double factor[2]) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s && factor, "null argument");
    double2* amps = s->arr[DFSA_AMPS];
    double2 c = hostAmp(factor);
    bool negate = (c.x == -1.0 && c.y == 0.0);
    auto ld = [=] __device__(uint64_t i) { return Amp1{amps[i]}; };
    auto st = [=] __device__(uint64_t i, const Amp1& v) { amps[i] = negate ? make_double2(-v.a.x, -v.a.y) : cmul(v.a, c); };
    return launchStream<2, Amp1>(s->numAmps, ld, st);
}

// ---------------------------------------------------------------------------------------------------------
// K8 / K10 / K19-K22 building blocks on a sub-cube: k = j with `values` bits inserted at sorted `positions`.
static int subcubeSpec(dfsa_state* s, const uint32_t* positions, unsigned n, uint64_t values, BitSpec* spec, uint64_t* fixed) {
    DFSA_REQUIRE(s && s->arr[DFSA_BUFFER], "null state / no exchange buffer");
    DFSA_REQUIRE(n >= 1 && n <= s->logNumAmps, "bad number of positions");
    for (unsigned q = 1; q < n; q++) DFSA_REQUIRE(positions[q] > positions[q - 1], "positions must be strictly increasing");
    DFSA_TRY(sortedSpec(positions, n, s->logNumAmps, spec, nullptr));
    *fixed = depositBits(values, *spec);
    return DFSA_OK;
}

extern "C" int dfsa_k_pack(dfsa_state* s, const uint32_t* positions, unsigned numPositions, uint64_t values, uint64_t dstStart) {
    DFSA_TRY(dfsaEnsureDevice());
    BitSpec spec; uint64_t fixed;
    DFSA_TRY(subcubeSpec(s, positions, numPositions, values, &spec, &fixed));
    uint64_t items = s->numAmps >> numPositions;
    DFSA_REQUIRE(dstStart + items <= s->numAmps, "pack overruns the buffer");
    const double2* amps = s->arr[DFSA_AMPS];
    double2* buf = s->arr[DFSA_BUFFER] + dstStart;
    auto ld = [=] __device__(uint64_t j) { return Amp1{amps[insertZeroBits(j, spec) | fixed]}; };
    auto st = [=] __device__(uint64_t j, const Amp1& v) { buf[j] = v.a; };
    return launchStream<2, Amp1>(items, ld, st);
}

extern "C" int dfsa_k_unpack(dfsa_state* s, const uint32_t* positions, unsigned numPositions, uint64_t values, uint64_t srcStart) {
    DFSA_TRY(dfsaEnsureDevice());
    BitSpec spec; uint64_t fixed;
    DFSA_TRY(subcubeSpec(s, positions, numPositions, values, &spec, &fixed));
    uint64_t items = s->numAmps >> numPositions;
    DFSA_REQUIRE(srcStart + items <= s->numAmps, "unpack overruns the buffer");
    double2* amps = s->arr[DFSA_AMPS];
    const double2* buf = s->arr[DFSA_BUFFER] + srcStart;
    auto ld = [=] __device__(uint64_t j) { return Amp1{buf[j]}; };
    auto st = [=] __device__(uint64_t j, const Amp1& v) { amps[insertZeroBits(j, spec) | fixed] = v.a; };
    return launchStream<2, Amp1>(items, ld, st);
}

int dfsaCombineSub(dfsa_state* s, const uint32_t* positions, unsigned numPositions, uint64_t values, uint64_t srcStart,
                   const double f0[2], const double f1[2], bool partnerFirst);
extern "C" int dfsa_k_combineSub(dfsa_state* s, const uint32_t* positions, unsigned numPositions, uint64_t values, uint64_t srcStart,
                                 const double f0[2], const double f1[2]) {
    return dfsaCombineSub(s, positions, numPositions, values, srcStart, f0, f1, false);
}
int dfsaCombineSub(dfsa_state* s, const uint32_t* positions, unsigned numPositions, uint64_t values, uint64_t srcStart,
                   const double f0[2], const double f1[2], bool partnerFirst) {
    DFSA_TRY(dfsaEnsureDevice());
    BitSpec spec; uint64_t fixed;
    DFSA_TRY(subcubeSpec(s, positions, numPositions, values, &spec, &fixed));
    uint64_t items = s->numAmps >> numPositions;
    DFSA_REQUIRE(srcStart + items <= s->numAmps && f0 && f1, "combineSub overruns the buffer / null factor");
    double2* amps = s->arr[DFSA_AMPS];
    const double2* buf = s->arr[DFSA_BUFFER] + srcStart;
    double2 c0 = hostAmp(f0), c1 = hostAmp(f1);
    auto ld = [=] __device__(uint64_t j) { return Amp2{amps[insertZeroBits(j, spec) | fixed], buf[j]}; };
    auto st = [=] __device__(uint64_t j, const Amp2& v) { amps[insertZeroBits(j, spec) | fixed] = partnerFirst ? cfma(c0, v.a0, cmul(c1, v.a1)) : cfma(c1, v.a1, cmul(c0, v.a0)); };
    return launchStream<1, Amp2>(items, ld, st);
}

// X2+K8 over peer memory (distributed_statevector.hpp:43-78: a 2x2 gate on a prefix target with suffix controls): only the
// sub-cube with every suffix control set is touched, so the result cannot go out of place without copying the whole shard.
// Two small passes instead: buffer[j] = f0*amps[k(j)] + f1*partner_amps[k(j)] (partner read over NVLink), then -- once both
// ranks have finished reading each other -- amps[k(j)] = buffer[j] (dfsa_k_unpack). 64 bytes of HBM traffic per touched
// amplitude against the staged path's 112 (pack, send + receive staging, combine), NVLink 16 per direction either way.
int dfsaLaunchFusedCombineSub(dfsa_state* s, const double2* remote, const BitSpec& spec, uint64_t fixed, double2 c0, double2 c1, bool partnerFirst) {
    const double2* amps = s->arr[DFSA_AMPS];
    double2* out = s->arr[DFSA_BUFFER];
    auto ld = [=] __device__(uint64_t j) { const uint64_t k = insertZeroBits(j, spec) | fixed; return Amp2{amps[k], remote[k]}; };
    auto st = [=] __device__(uint64_t j, const Amp2& v) { out[j] = partnerFirst ? cfma(c0, v.a0, cmul(c1, v.a1)) : cfma(c1, v.a1, cmul(c0, v.a0)); };
    return launchStreamRemote<Amp2>(s->numAmps >> spec.n, ld, st);
}

// unpack half a shard from an arbitrary (possibly peer-mapped) source: amps[insert(k, qb, bitValue)] = src[k]
// Suffix<->prefix qubit swap (distributed_statevector.hpp:140-186) as ONE out-of-place pass over peer memory:
//   buffer[j] = amps[j]                     where bit qb of j equals this rank's bit of the prefix qubit (stays)
//   buffer[j] = partner_amps[j ^ (1 << qb)] otherwise (the partner's moving half, read over NVLink)
// then the caller swaps amps <-> buffer. No pack pass, no staging copy; item k touches amps[i], remote[i] with
// i = k with myBit inserted at qb. Pure moves: bit-exact.
int dfsaLaunchFusedSwap(dfsa_state* s, const double2* remote, unsigned qb, unsigned myBit) {
    const double2* amps = s->arr[DFSA_AMPS];
    double2* out = s->arr[DFSA_BUFFER];
    const uint64_t bit = 1ULL << qb, fixed = (uint64_t)(myBit & 1u) << qb;
    auto ld = [=] __device__(uint64_t k) { const uint64_t i = insertZeroBit(k, qb) | fixed; return Amp2{amps[i], remote[i]}; };
    auto st = [=] __device__(uint64_t k, const Amp2& v) {
        const uint64_t i = insertZeroBit(k, qb) | fixed;
        out[i] = v.a0;
        out[i ^ bit] = v.a1;
    };
    return launchStreamRemote<Amp2>(s->numAmps >> 1, ld, st);
}

// Single-shot relocation: swap k suffix qubits s_i with k prefix qubits in ONE out-of-place pass over the
// 2^k peer shards of this rank's group, instead of k sequential suffix<->prefix swaps (distributed_statevector.hpp:213-223).
//   buffer[j] = shard_of_rank(R with its swapped rank bits := the s-bits of j)[ j with its s-bits := R's swapped rank bits ]
// `peers[sigma]` is that shard for s-bits sigma (this rank's own amps for sigma == rho). Pure moves: bit-exact.
struct RelocPeers { const double2* shard[16]; };
struct RelocBits { unsigned pos[4]; };

int dfsaLaunchRelocate(dfsa_state* s, const double2* const* peers, const unsigned* suffixPos, unsigned k, unsigned rho) {
    RelocPeers table;
    RelocBits bits;
    uint64_t sMask = 0, rhoBits = 0;
    for (unsigned i = 0; i < 4; i++) bits.pos[i] = i < k ? suffixPos[i] : 0u;
    for (unsigned i = 0; i < k; i++) { sMask |= 1ULL << suffixPos[i]; rhoBits |= (uint64_t)((rho >> i) & 1u) << suffixPos[i]; }
    for (unsigned g = 0; g < 16; g++) table.shard[g] = g < (1u << k) ? peers[g] : nullptr;
    double2* out = s->arr[DFSA_BUFFER];
    // Item jj stands for output index j = jj ^ rhoBits: every rank walks the 2^k source shards starting with its OWN one. Walking
    // them in the same order on every rank makes all 2^k ranks of a group pull from the same owner at the same time (its NVLink
    // egress shared 2^k - 1 ways while the other owners idle): 72.8 ms instead of 19 for the 2-pair relocation of partialTrace at
    // 4 GPUs (profiles/r02_bench_n4.json). With the XOR, at any time the readers of a group form a perfect matching.
    auto ld = [=] __device__(uint64_t jj) {
        const uint64_t j = jj ^ rhoBits;
        unsigned sigma = 0;
        for (unsigned i = 0; i < k; i++) sigma |= (unsigned)((j >> bits.pos[i]) & 1ULL) << i;
        return Amp1{table.shard[sigma][(j & ~sMask) | rhoBits]};
    };
    auto st = [=] __device__(uint64_t jj, const Amp1& v) { out[jj ^ rhoBits] = v.a; };
    return launchStreamRemote<Amp1>(s->numAmps, ld, st);
}

// K9: distributed_statevector.hpp:133-135, 152-156
extern "C" int dfsa_k_copyFromBuffer(dfsa_state* s, uint64_t dstStart, uint64_t srcStart, uint64_t num) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(s && s->arr[DFSA_BUFFER], "null state / no exchange buffer");
    DFSA_REQUIRE(dstStart + num <= s->numAmps && srcStart + num <= s->numAmps, "copy out of range");
    if (num == s->numAmps) {
        // whole shard: the received buffer simply BECOMES the shard (no 32*A-byte copy). Kernels already enqueued hold
        // the old pointers by value; everything enqueued later sees the new ones; peers look the current registry
        // slots of this state up in the shared page (dfsaPublishArrays).
        return dfsa_state_swap_arrays(s);
    }
    DFSA_CUDA(cudaMemcpyAsync(s->arr[DFSA_AMPS] + dstStart, s->arr[DFSA_BUFFER] + srcStart, num * sizeof(double2), cudaMemcpyDeviceToDevice, dfsaCtx().compute));
    return DFSA_OK;
}


double2 dfsaPowIHost(unsigned k) { return powIHost(k); }
