// dfsa_kernels_dm.cu -- density-matrix kernels K12-K17, K19-K23 for sm_100a.
// The density matrix is a Choi vector of 2N index bits: bit q (< N) is the ket/row bit of qubit q, bit q+N its
// bra/column bit; flat = 2^N*col + row (tests/test_utilities.hpp:498-500). Ranks own column blocks, so the
// "prefix" qubits are the bra bits q+N >= logNumAmps, i.e. q >= N - log2(P).
#include <algorithm>
#include <vector>

#include <string.h>
#include "dfsa_stream_kernels.cuh"

namespace {
inline unsigned prefixThreshold(const dfsa_state* s) { return s->numQubits - s->logNumNodes; }
inline int rankBit(const dfsa_state* s, unsigned qb) { return (s->rank >> (qb - prefixThreshold(s))) & 1; }
}  // namespace

#define DFSA_DM_ENTRY(s)                              \
    DFSA_TRY(dfsaEnsureDevice());                     \
    DFSA_REQUIRE((s) && (s)->isDensity, "needs a density-matrix state")

// ---------------------------------------------------------------------------------------------------------
// K12: local_densitymatrix.hpp:12-42. Scale the amplitudes whose ket bit != bra bit by 1-2p. 16*A bytes.
extern "C" int dfsa_k_oneQubitDephasing(dfsa_state* s, unsigned qb, double prob) {
    DFSA_DM_ENTRY(s);
    DFSA_REQUIRE(qb < s->numQubits, "qubit out of range");
    const double fac = 1 - 2 * prob;
    double2* amps = s->arr[DFSA_AMPS];
    if (qb >= prefixThreshold(s)) {
        // bra bit lives in the rank index: scale the half whose ket bit differs from it
        const uint64_t fixed = (uint64_t)(!rankBit(s, qb)) << qb;
        auto ld = [=] __device__(uint64_t k) { return Amp1{amps[insertZeroBit(k, qb) | fixed]}; };
        auto st = [=] __device__(uint64_t k, const Amp1& v) { amps[insertZeroBit(k, qb) | fixed] = cscale(fac, v.a); };
        return launchStream<2, Amp1>(s->numAmps >> 1, ld, st);
    }
    const unsigned alt = qb + s->numQubits;
    const uint64_t bKet = 1ULL << qb, bBra = 1ULL << alt;
    auto ld = [=] __device__(uint64_t k) {
        uint64_t j00 = insertZeroBit(insertZeroBit(k, qb), alt);
        return Amp2{amps[j00 | bKet], amps[j00 | bBra]};
    };
    auto st = [=] __device__(uint64_t k, const Amp2& v) {
        uint64_t j00 = insertZeroBit(insertZeroBit(k, qb), alt);
        amps[j00 | bKet] = cscale(fac, v.a0);
        amps[j00 | bBra] = cscale(fac, v.a1);
    };
    return launchStream<1, Amp2>(s->numAmps >> 2, ld, st);
}

// K13: local_densitymatrix.hpp:45-60. amps[j] *= 1 - 4p/3 where either qubit's ket/bra bits differ (global index).
extern "C" int dfsa_k_twoQubitDephasing(dfsa_state* s, unsigned qb1, unsigned qb2, double prob) {
    DFSA_DM_ENTRY(s);
    DFSA_REQUIRE(qb1 < s->numQubits && qb2 < s->numQubits && qb1 != qb2, "needs two distinct qubits");
    const unsigned N = s->numQubits;
    const double fac = -4 * prob / 3 + 1.;
    const uint64_t rs = (uint64_t)s->rank << s->logNumAmps;
    double2* amps = s->arr[DFSA_AMPS];
    auto ld = [=] __device__(uint64_t j) { return Amp1{amps[j]}; };
    auto st = [=] __device__(uint64_t j, const Amp1& v) {
        uint64_t i = rs | j;
        unsigned differ = (unsigned)(((i >> qb1) ^ (i >> (qb1 + N))) | ((i >> qb2) ^ (i >> (qb2 + N)))) & 1u;
        if (differ) amps[j] = cscale(fac, v.a);            // untouched amplitudes are not written back
    };
    return launchStream<2, Amp1>(s->numAmps, ld, st);
}

// K14: local_densitymatrix.hpp:63-81 (suffix qubit). c1 = 2p/3, c2 = 1-2p/3, c3 = 1-4p/3
// (distributed_densitymatrix.hpp:106-108). 32*A bytes.
extern "C" int dfsa_k_oneQubitDepolarising(dfsa_state* s, unsigned qb, double prob) {
    DFSA_DM_ENTRY(s);
    DFSA_REQUIRE(qb < prefixThreshold(s), "local kernel needs a suffix qubit");
    const double c1 = 2 * prob / 3, c2 = 1 - 2 * prob / 3, c3 = 1 - 4 * prob / 3;
    const unsigned alt = qb + s->numQubits;
    const uint64_t bKet = 1ULL << qb, bBra = 1ULL << alt;
    double2* amps = s->arr[DFSA_AMPS];
    auto ld = [=] __device__(uint64_t k) {
        uint64_t j00 = insertZeroBit(insertZeroBit(k, qb), alt);
        return Amp4{amps[j00], amps[j00 | bKet], amps[j00 | bBra], amps[j00 | bKet | bBra]};
    };
    auto st = [=] __device__(uint64_t k, const Amp4& v) {
        uint64_t j00 = insertZeroBit(insertZeroBit(k, qb), alt);
        amps[j00]               = make_double2(fma(c1, v.a11.x, c2 * v.a00.x), fma(c1, v.a11.y, c2 * v.a00.y));
        amps[j00 | bKet]        = cscale(c3, v.a01);
        amps[j00 | bBra]        = cscale(c3, v.a10);
        amps[j00 | bKet | bBra] = make_double2(fma(c2, v.a11.x, c1 * v.a00.x), fma(c2, v.a11.y, c1 * v.a00.y));
    };
    return launchStream<1, Amp4>(s->numAmps >> 2, ld, st);
}

// K16: local_densitymatrix.hpp:111-131 (suffix qubit). a00 += p*a11; a01,a10 *= sqrt(1-p); a11 *= 1-p.
extern "C" int dfsa_k_damping(dfsa_state* s, unsigned qb, double prob) {
    DFSA_DM_ENTRY(s);
    DFSA_REQUIRE(qb < prefixThreshold(s), "local kernel needs a suffix qubit");
    const double c1 = sqrt(1 - prob), c2 = 1 - prob;
    const unsigned alt = qb + s->numQubits;
    const uint64_t bKet = 1ULL << qb, bBra = 1ULL << alt;
    double2* amps = s->arr[DFSA_AMPS];
    auto ld = [=] __device__(uint64_t k) {
        uint64_t j00 = insertZeroBit(insertZeroBit(k, qb), alt);
        return Amp4{amps[j00], amps[j00 | bKet], amps[j00 | bBra], amps[j00 | bKet | bBra]};
    };
    auto st = [=] __device__(uint64_t k, const Amp4& v) {
        uint64_t j00 = insertZeroBit(insertZeroBit(k, qb), alt);
        amps[j00]               = make_double2(fma(prob, v.a11.x, v.a00.x), fma(prob, v.a11.y, v.a00.y));
        amps[j00 | bKet]        = cscale(c1, v.a01);
        amps[j00 | bBra]        = cscale(c1, v.a10);
        amps[j00 | bKet | bBra] = cscale(c2, v.a11);
    };
    return launchStream<1, Amp4>(s->numAmps >> 2, ld, st);
}

// K15: local_densitymatrix.hpp:83-108 (both qubits suffix). Pass 1 scales every amplitude whose (ket,bra) bits
// of either qubit differ by 1+c3; pass 2 mixes the four "diagonal" amplitudes of each 16-group:
//   a' = c1*a + c2*(a0000 + a0101 + a1010 + a1111)   (sum in that order).
// Reference constants (distributed_densitymatrix.hpp:247-249): c1 = 1-4p/5, c2 = 4p/15, c3 = -16p/15 -- with
// these the map is NOT the depolarising channel but it is what the reference computes; `corrected`
// uses c1 = 1-16p/15, which is the channel (1-16p/15) rho + (4p/15) I (x) Tr_2 rho.
// ONE pass (32*A bytes; the reference makes two, 40*A): an item is the 16 amplitudes spanned by the four bits
// (q0, q1 ket; q2, q3 bra); the 12 off-"diagonal" ones are scaled, the 4 diagonal ones mixed. Each item is read whole
// before it is written, and pass 1 never touches what pass 2 reads, so the arithmetic is the reference's, operation for
// operation.
struct Amp16 { static constexpr int kLoads = 16; double2 v[16]; };

extern "C" int dfsa_k_twoQubitDepolarising(dfsa_state* s, unsigned qb1, unsigned qb2, double prob, int corrected) {
    DFSA_DM_ENTRY(s);
    if (qb1 > qb2) std::swap(qb1, qb2);
    DFSA_REQUIRE(qb1 != qb2 && qb2 < prefixThreshold(s), "local kernel needs two distinct suffix qubits");
    const unsigned N = s->numQubits, q0 = qb1, q1 = qb2, q2 = qb1 + N, q3 = qb2 + N;
    const double c1 = corrected ? 1 - 16 * prob / 15 : 1 - 4 * prob / 5, c2 = 4 * prob / 15, c3 = -16 * prob / 15;
    const double offFac = 1. + c3;
    double2* amps = s->arr[DFSA_AMPS];
    // group element e = (b3 b2 b1 b0): bit i of e sits at index bit q_i; diagonal <=> b0 == b2 and b1 == b3: e = 0, 5, 10, 15
    auto index = [=] __device__(uint64_t k, unsigned e) {
        uint64_t j = insertZeroBit(insertZeroBit(insertZeroBit(insertZeroBit(k, q0), q1), q2), q3);
        return j | ((uint64_t)(e & 1u) << q0) | ((uint64_t)((e >> 1) & 1u) << q1) | ((uint64_t)((e >> 2) & 1u) << q2) | ((uint64_t)((e >> 3) & 1u) << q3);
    };
    auto ld = [=] __device__(uint64_t k) {
        Amp16 it;
#pragma unroll
        for (unsigned e = 0; e < 16; e++) it.v[e] = amps[index(k, e)];
        return it;
    };
    auto st = [=] __device__(uint64_t k, const Amp16& it) {
        const double2 a00 = it.v[0], a01 = it.v[5], a10 = it.v[10], a11 = it.v[15];
        const double tx = ((a00.x + a01.x) + a10.x) + a11.x, ty = ((a00.y + a01.y) + a10.y) + a11.y;
#pragma unroll
        for (unsigned e = 0; e < 16; e++) {
            const bool diag = (e == 0 || e == 5 || e == 10 || e == 15);
            amps[index(k, e)] = diag ? make_double2(fma(c2, tx, c1 * it.v[e].x), fma(c2, ty, c1 * it.v[e].y)) : cscale(offFac, it.v[e]);
        }
    };
    return launchStream<1, Amp16>(s->numAmps >> 4, ld, st);
}

// K19: distributed_densitymatrix.hpp:130-141. After the half exchange (received half at buffer[A/2..)):
//   other half (ket bit != rank bit) *= c3;   this half: a = c2*a + c1*recv.
extern "C" int dfsa_k_depol1Combine(dfsa_state* s, unsigned qb, unsigned bit, double prob) {
    DFSA_DM_ENTRY(s);
    DFSA_REQUIRE(qb >= prefixThreshold(s) && qb < s->numQubits && s->arr[DFSA_BUFFER], "needs a prefix qubit and the exchange buffer");
    const double c1 = 2 * prob / 3, c2 = 1 - 2 * prob / 3, c3 = 1 - 4 * prob / 3;
    const uint64_t half = s->numAmps >> 1, same = (uint64_t)(bit & 1u) << qb, other = (uint64_t)(!(bit & 1u)) << qb;
    double2* amps = s->arr[DFSA_AMPS];
    const double2* recv = s->arr[DFSA_BUFFER] + half;
    using Item = Amp3;     // a = mine, b = other half, c = received
    auto ld = [=] __device__(uint64_t k) {
        uint64_t j = insertZeroBit(k, qb);
        return Item{amps[j | same], amps[j | other], recv[k]};
    };
    auto st = [=] __device__(uint64_t k, const Item& v) {
        uint64_t j = insertZeroBit(k, qb);
        amps[j | other] = cscale(c3, v.b);
        amps[j | same]  = make_double2(fma(c1, v.c.x, c2 * v.a.x), fma(c1, v.c.y, c2 * v.a.y));
    };
    return launchStream<1, Item>(half, ld, st);
}

// K19 fused with its exchange (distributed_densitymatrix.hpp:119-141 as ONE out-of-place pass over peer memory):
//   buffer[j | same]  = c2 * amps[j | same] + c1 * partner_amps[j | other]      (same = ket bit == this rank's bra bit)
//   buffer[j | other] = c3 * amps[j | other]
// `partner_amps[j | other]` is exactly what the reference packs, sends and receives; the caller swaps amps <-> buffer.
int dfsaLaunchFusedDepol1(dfsa_state* s, const double2* remote, unsigned qb, unsigned bit, double prob) {
    const double c1 = 2 * prob / 3, c2 = 1 - 2 * prob / 3, c3 = 1 - 4 * prob / 3;
    const uint64_t same = (uint64_t)(bit & 1u) << qb, other = (uint64_t)(!(bit & 1u)) << qb;
    const double2* amps = s->arr[DFSA_AMPS];
    double2* out = s->arr[DFSA_BUFFER];
    using Item = Amp3;     // a = mine, b = other half, c = partner's
    auto ld = [=] __device__(uint64_t k) {
        uint64_t j = insertZeroBit(k, qb);
        return Item{amps[j | same], amps[j | other], remote[j | other]};
    };
    auto st = [=] __device__(uint64_t k, const Item& v) {
        uint64_t j = insertZeroBit(k, qb);
        out[j | other] = cscale(c3, v.b);
        out[j | same]  = make_double2(fma(c1, v.c.x, c2 * v.a.x), fma(c1, v.c.y, c2 * v.a.y));
    };
    return launchStreamRemote<Item>(s->numAmps >> 1, ld, st);
}

// K22 fused with its one-way transfer (distributed_densitymatrix.hpp:284-313), out of place:
//   bra bit 1 ranks: buffer[ket 1] = (1-p) amps[ket 1],  buffer[ket 0] = sqrt(1-p) amps[ket 0]            (no remote read)
//   bra bit 0 ranks: buffer[ket 0] = amps[ket 0] + p * partner_amps[ket 1],  buffer[ket 1] = sqrt(1-p) amps[ket 1]
int dfsaLaunchFusedDamping(dfsa_state* s, const double2* remote, unsigned qb, unsigned bit, double prob) {
    const double c1 = sqrt(1 - prob), c2 = 1 - prob;
    const uint64_t one = 1ULL << qb;
    const double2* amps = s->arr[DFSA_AMPS];
    double2* out = s->arr[DFSA_BUFFER];
    if (bit & 1u) {
        auto ld = [=] __device__(uint64_t k) { uint64_t j = insertZeroBit(k, qb); return Amp2{amps[j], amps[j | one]}; };
        auto st = [=] __device__(uint64_t k, const Amp2& v) {
            uint64_t j = insertZeroBit(k, qb);
            out[j] = cscale(c1, v.a0);
            out[j | one] = cscale(c2, v.a1);
        };
        return launchStream<1, Amp2>(s->numAmps >> 1, ld, st);
    }
    auto ld = [=] __device__(uint64_t k) { uint64_t j = insertZeroBit(k, qb); return Amp3{amps[j], amps[j | one], remote[j | one]}; };
    auto st = [=] __device__(uint64_t k, const Amp3& v) {
        uint64_t j = insertZeroBit(k, qb);
        out[j] = make_double2(fma(prob, v.c.x, v.a.x), fma(prob, v.c.y, v.a.y));
        out[j | one] = cscale(c1, v.b);
    };
    return launchStreamRemote<Amp3>(s->numAmps >> 1, ld, st);
}

// K20: distributed_densitymatrix.hpp:152-183 (qb1 suffix, qb2 prefix). q0 = qb1, q1 = qb2, q2 = qb1+N, bit = rank bit of qb2's bra.
// phase 0: scale + pack the pre-summed eighth into buffer[0..A/8); phase 1: combine with buffer[A/8..A/4).
// `phase | DFSA_DEPOL2_CORRECTED` selects the true channel (1-16p/15) rho + (4p/15) I (x) Tr_2 rho instead of the reference's
// formulas: new = c1' a + c2 ((a0b0 + a1b1) + received), with the OLD a0b0 in both lines (no read-after-write).
extern "C" int dfsa_k_depol2Pair(dfsa_state* s, unsigned q0, unsigned q1, unsigned q2, unsigned bit, double prob, int phase) {
    DFSA_DM_ENTRY(s);
    DFSA_REQUIRE(q0 < q1 && q1 < q2 && q2 < s->logNumAmps && s->arr[DFSA_BUFFER], "bad qubits / no exchange buffer");
    const bool corrected = (phase & DFSA_DEPOL2_CORRECTED) != 0;
    phase &= ~DFSA_DEPOL2_CORRECTED;
    const double c1 = corrected ? 1 - 16 * prob / 15 : 1 - 4 * prob / 5, c2 = 4 * prob / 15, c3 = -16 * prob / 15;
    const uint64_t eighth = s->numAmps >> 3;
    const uint64_t b1 = (uint64_t)(bit & 1u) << q1, b02 = (1ULL << q0) | (1ULL << q2);
    double2* amps = s->arr[DFSA_AMPS];
    double2* buf = s->arr[DFSA_BUFFER];
    if (phase == 0) {
        const double offFac = 1. + c3;
        auto ld = [=] __device__(uint64_t j) { return Amp1{amps[j]}; };
        auto st = [=] __device__(uint64_t j, const Amp1& v) {
            unsigned f1 = !(((j >> q0) ^ (j >> q2)) & 1ULL), f2 = (((j >> q1) & 1ULL) == (bit & 1u));
            if (!(f1 & f2)) amps[j] = cscale(offFac, v.a);
        };
        DFSA_TRY((launchStream<2, Amp1>(s->numAmps, ld, st)));
        auto ld2 = [=] __device__(uint64_t k) {
            uint64_t j0b0 = insertZeroBit(insertZeroBit(insertZeroBit(k, q0), q1), q2) | b1;
            return Amp2{amps[j0b0], amps[j0b0 | b02]};
        };
        auto st2 = [=] __device__(uint64_t k, const Amp2& v) { buf[k] = cadd(v.a0, v.a1); };
        return launchStream<1, Amp2>(eighth, ld2, st2);
    }
    DFSA_REQUIRE(phase == 1, "phase must be 0 or 1");
    using Item = Amp3;     // a = a0b0, b = a1b1, c = received
    auto ld = [=] __device__(uint64_t k) {
        uint64_t j0b0 = insertZeroBit(insertZeroBit(insertZeroBit(k, q0), q1), q2) | b1;
        return Item{amps[j0b0], amps[j0b0 | b02], buf[k + eighth]};
    };
    auto st = [=] __device__(uint64_t k, const Item& v) {
        uint64_t j0b0 = insertZeroBit(insertZeroBit(insertZeroBit(k, q0), q1), q2) | b1;
        double2 n0, n1;
        if (corrected) {
            const double sx = (v.a.x + v.b.x) + v.c.x, sy = (v.a.y + v.b.y) + v.c.y;
            n0 = make_double2(fma(c2, sx, c1 * v.a.x), fma(c2, sy, c1 * v.a.y));
            n1 = make_double2(fma(c2, sx, c1 * v.b.x), fma(c2, sy, c1 * v.b.y));
        } else {
            // literal reference arithmetic, including the read-after-write of :181 -> :182
            n0 = make_double2(fma(c2, v.b.x + v.c.x, c1 * v.a.x), fma(c2, v.b.y + v.c.y, c1 * v.a.y));
            n1 = make_double2(fma(c2, n0.x + v.c.x, c1 * v.b.x), fma(c2, n0.y + v.c.y, c1 * v.b.y));
        }
        amps[j0b0] = n0;
        amps[j0b0 | b02] = n1;
    };
    return launchStream<1, Item>(eighth, ld, st);
}

// K21: distributed_densitymatrix.hpp:195-237 (both qubits prefix).
// phase 0: scale + pack quarter; phase 1: a = c1*a + c2*recv, also repacked; phase 2: a = (c2/c1)*recv  (overwrite, as the reference)
// `phase | DFSA_DEPOL2_CORRECTED`: the true channel -- phase 1 keeps amps and sends on t = a + received, phase 2 computes
// a = c1' a + c2 (t + received) (t is still in buffer[0..A/4)).
extern "C" int dfsa_k_depol2Quad(dfsa_state* s, unsigned q0, unsigned q1, unsigned bit0, unsigned bit1, double prob, int phase) {
    DFSA_DM_ENTRY(s);
    DFSA_REQUIRE(q0 < q1 && q1 < s->logNumAmps && s->arr[DFSA_BUFFER], "bad qubits / no exchange buffer");
    const bool corrected = (phase & DFSA_DEPOL2_CORRECTED) != 0;
    phase &= ~DFSA_DEPOL2_CORRECTED;
    const double c1 = corrected ? 1 - 16 * prob / 15 : 1 - 4 * prob / 5, c2 = 4 * prob / 15, c3 = -16 * prob / 15;
    const uint64_t quarter = s->numAmps >> 2;
    const uint64_t fixed = ((uint64_t)(bit0 & 1u) << q0) | ((uint64_t)(bit1 & 1u) << q1);
    double2* amps = s->arr[DFSA_AMPS];
    double2* buf = s->arr[DFSA_BUFFER];
    if (phase == 0) {
        const double offFac = 1. + c3;
        auto ld = [=] __device__(uint64_t j) { return Amp1{amps[j]}; };
        auto st = [=] __device__(uint64_t j, const Amp1& v) {
            unsigned f1 = (((j >> q0) & 1ULL) == (bit0 & 1u)), f2 = (((j >> q1) & 1ULL) == (bit1 & 1u));
            if (!(f1 & f2)) amps[j] = cscale(offFac, v.a);
        };
        DFSA_TRY((launchStream<2, Amp1>(s->numAmps, ld, st)));
        auto ld2 = [=] __device__(uint64_t k) { return Amp1{amps[insertZeroBit(insertZeroBit(k, q0), q1) | fixed]}; };
        auto st2 = [=] __device__(uint64_t k, const Amp1& v) { buf[k] = v.a; };
        return launchStream<2, Amp1>(quarter, ld2, st2);
    }
    if (phase == 1) {
        auto ld = [=] __device__(uint64_t k) { return Amp2{amps[insertZeroBit(insertZeroBit(k, q0), q1) | fixed], buf[k + quarter]}; };
        auto st = [=] __device__(uint64_t k, const Amp2& v) {
            if (corrected) { buf[k] = cadd(v.a0, v.a1); return; }
            double2 n = make_double2(fma(c2, v.a1.x, c1 * v.a0.x), fma(c2, v.a1.y, c1 * v.a0.y));
            amps[insertZeroBit(insertZeroBit(k, q0), q1) | fixed] = n;
            buf[k] = n;
        };
        return launchStream<1, Amp2>(quarter, ld, st);
    }
    DFSA_REQUIRE(phase == 2, "phase must be 0, 1 or 2");
    if (corrected) {
        auto ld = [=] __device__(uint64_t k) { return Amp3{amps[insertZeroBit(insertZeroBit(k, q0), q1) | fixed], buf[k], buf[k + quarter]}; };
        auto st = [=] __device__(uint64_t k, const Amp3& v) {
            const double sx = v.b.x + v.c.x, sy = v.b.y + v.c.y;
            amps[insertZeroBit(insertZeroBit(k, q0), q1) | fixed] = make_double2(fma(c2, sx, c1 * v.a.x), fma(c2, sy, c1 * v.a.y));
        };
        return launchStream<1, Amp3>(quarter, ld, st);
    }
    const double c4 = c2 / c1;
    auto ld = [=] __device__(uint64_t k) { return Amp1{buf[k + quarter]}; };
    auto st = [=] __device__(uint64_t k, const Amp1& v) { amps[insertZeroBit(insertZeroBit(k, q0), q1) | fixed] = cscale(c4, v.a); };
    return launchStream<2, Amp1>(quarter, ld, st);
}

// K20 fused with its exchange (distributed_densitymatrix.hpp:152-183 as ONE out-of-place pass over peer memory). An item is
// the 8 amplitudes spanned by (q0, q1 ket bits; q2 = bra bit of qb1); the two "diagonal" ones with q1 bit == this rank's
// bra bit of qb2 (a0b0, a1b1) are mixed with the partner's pair, which the reference pre-sums, packs and sends -- here the
// partner's two amplitudes are read over NVLink and summed in the same order; everything else is scaled. The partner's
// diagonal amplitudes are untouched by its own scaling step, so reading its ORIGINAL shard gives the reference's values.
struct Amp10 { static constexpr int kLoads = 10; double2 v[8]; double2 r0, r1; };

int dfsaLaunchFusedDepol2Pair(dfsa_state* s, const double2* remote, unsigned q0, unsigned q1, unsigned q2, unsigned bit, double prob, bool corrected) {
    const double c1 = corrected ? 1 - 16 * prob / 15 : 1 - 4 * prob / 5, c2 = 4 * prob / 15, offFac = 1. + (-16 * prob / 15);
    const double2* amps = s->arr[DFSA_AMPS];
    double2* out = s->arr[DFSA_BUFFER];
    const unsigned mine = bit & 1u;
    const uint64_t other1 = (uint64_t)(mine ^ 1u) << q1, b02 = (1ULL << q0) | (1ULL << q2);
    auto index = [=] __device__(uint64_t k, unsigned e) {     // e = (z y x): x at q0, y at q1, z at q2
        return insertZeroBit(insertZeroBit(insertZeroBit(k, q0), q1), q2) | ((uint64_t)(e & 1u) << q0) | ((uint64_t)((e >> 1) & 1u) << q1) | ((uint64_t)((e >> 2) & 1u) << q2);
    };
    auto ld = [=] __device__(uint64_t k) {
        Amp10 it;
#pragma unroll
        for (unsigned e = 0; e < 8; e++) it.v[e] = amps[index(k, e)];
        const uint64_t jz = index(k, 0) | other1;
        it.r0 = remote[jz];
        it.r1 = remote[jz | b02];
        return it;
    };
    auto st = [=] __device__(uint64_t k, const Amp10& it) {
        const unsigned eA = mine << 1, eB = eA | 5u;             // a0b0: x = z = 0, y = mine;  a1b1: x = z = 1
        const double2 A = mine ? it.v[2] : it.v[0], B = mine ? it.v[7] : it.v[5], recv = cadd(it.r0, it.r1);
        double2 n0, n1;
        if (corrected) {
            const double sx = (A.x + B.x) + recv.x, sy = (A.y + B.y) + recv.y;
            n0 = make_double2(fma(c2, sx, c1 * A.x), fma(c2, sy, c1 * A.y));
            n1 = make_double2(fma(c2, sx, c1 * B.x), fma(c2, sy, c1 * B.y));
        } else {                                                 // the reference's lines :181-182, read-after-write included
            n0 = make_double2(fma(c2, B.x + recv.x, c1 * A.x), fma(c2, B.y + recv.y, c1 * A.y));
            n1 = make_double2(fma(c2, n0.x + recv.x, c1 * B.x), fma(c2, n0.y + recv.y, c1 * B.y));
        }
#pragma unroll
        for (unsigned e = 0; e < 8; e++) out[index(k, e)] = (e == eA) ? n0 : ((e == eB) ? n1 : cscale(offFac, it.v[e]));
    };
    return launchStream<1, Amp10, decltype(ld), decltype(st), true>(s->numAmps >> 3, ld, st);
}

// K21 fused with its two exchanges (distributed_densitymatrix.hpp:195-237), ONE out-of-place pass over the shards of the
// four ranks that differ in the two bra bits. An item is the 4 amplitudes spanned by the ket bits (q0, q1); the one whose
// ket bits equal this rank's bra bits is the "diagonal" one. The reference's net effect, literally: after the scaling step
//   diagonal = (c2/c1) * (c1 * P1 + c2 * P01)     P1 = partner 1's original diagonal amplitude, P01 = its partner 0's
// (the value of phase 1, c1 a + c2 P0, is overwritten by phase 2 at :236 -- SURVEY F2). Corrected: c1' a + c2 (((a + P0) + P1) + P01).
struct Amp7 { static constexpr int kLoads = 7; double2 v[4]; double2 p0, p1, p01; };

int dfsaLaunchFusedDepol2Quad(dfsa_state* s, const double2* const* remote /* P0, P1, P01 */, unsigned q0, unsigned q1, unsigned bit0, unsigned bit1, double prob, bool corrected) {
    const double c1 = corrected ? 1 - 16 * prob / 15 : 1 - 4 * prob / 5, c2 = 4 * prob / 15, offFac = 1. + (-16 * prob / 15), c4 = c2 / c1;
    const double2* amps = s->arr[DFSA_AMPS];
    const double2 *r0 = remote[0], *r1 = remote[1], *r01 = remote[2];
    double2* out = s->arr[DFSA_BUFFER];
    const unsigned b0 = bit0 & 1u, b1 = bit1 & 1u, eMine = b0 | (b1 << 1);
    const uint64_t f0 = ((uint64_t)(b0 ^ 1u) << q0) | ((uint64_t)b1 << q1), f1 = ((uint64_t)b0 << q0) | ((uint64_t)(b1 ^ 1u) << q1),
                   f01 = ((uint64_t)(b0 ^ 1u) << q0) | ((uint64_t)(b1 ^ 1u) << q1);
    auto ld = [=] __device__(uint64_t k) {
        Amp7 it;
        const uint64_t jz = insertZeroBit(insertZeroBit(k, q0), q1);
#pragma unroll
        for (unsigned e = 0; e < 4; e++) it.v[e] = amps[jz | ((uint64_t)(e & 1u) << q0) | ((uint64_t)(e >> 1) << q1)];
        it.p0 = corrected ? r0[jz | f0] : make_double2(0.0, 0.0);
        it.p1 = r1[jz | f1];
        it.p01 = r01[jz | f01];
        return it;
    };
    auto st = [=] __device__(uint64_t k, const Amp7& it) {
        const uint64_t jz = insertZeroBit(insertZeroBit(k, q0), q1);
        const double2 a = it.v[eMine];
        double2 n;
        if (corrected) {
            const double sx = ((a.x + it.p0.x) + it.p1.x) + it.p01.x, sy = ((a.y + it.p0.y) + it.p1.y) + it.p01.y;
            n = make_double2(fma(c2, sx, c1 * a.x), fma(c2, sy, c1 * a.y));
        } else {
            n = cscale(c4, make_double2(fma(c2, it.p01.x, c1 * it.p1.x), fma(c2, it.p01.y, c1 * it.p1.y)));
        }
#pragma unroll
        for (unsigned e = 0; e < 4; e++) out[jz | ((uint64_t)(e & 1u) << q0) | ((uint64_t)(e >> 1) << q1)] = (e == eMine) ? n : cscale(offFac, it.v[e]);
    };
    return launchStream<1, Amp7, decltype(ld), decltype(st), true>(s->numAmps >> 2, ld, st);
}

// K22: distributed_densitymatrix.hpp:284-313.
extern "C" int dfsa_k_dampingPrefix(dfsa_state* s, unsigned qb, unsigned bit, double prob, int phase) {
    DFSA_DM_ENTRY(s);
    DFSA_REQUIRE(qb >= prefixThreshold(s) && qb < s->numQubits && s->arr[DFSA_BUFFER], "needs a prefix qubit and the exchange buffer");
    const double c1 = sqrt(1 - prob), c2 = 1 - prob;
    const uint64_t half = s->numAmps >> 1;
    double2* amps = s->arr[DFSA_AMPS];
    double2* buf = s->arr[DFSA_BUFFER];
    if (phase == 0) {           // bit = 1 ranks: pack the ket=1 half, scale it by 1-p
        const uint64_t one = 1ULL << qb;
        auto ld = [=] __device__(uint64_t k) { return Amp1{amps[insertZeroBit(k, qb) | one]}; };
        auto st = [=] __device__(uint64_t k, const Amp1& v) { buf[k] = v.a; amps[insertZeroBit(k, qb) | one] = cscale(c2, v.a); };
        return launchStream<2, Amp1>(half, ld, st);
    }
    if (phase == 1) {           // every rank: the half with ket bit != rank bit decays by sqrt(1-p)
        const uint64_t fixed = (uint64_t)(!(bit & 1u)) << qb;
        auto ld = [=] __device__(uint64_t k) { return Amp1{amps[insertZeroBit(k, qb) | fixed]}; };
        auto st = [=] __device__(uint64_t k, const Amp1& v) { amps[insertZeroBit(k, qb) | fixed] = cscale(c1, v.a); };
        return launchStream<2, Amp1>(half, ld, st);
    }
    DFSA_REQUIRE(phase == 2, "phase must be 0, 1 or 2");   // bit = 0 ranks: a00 += p * a11 (received)
    auto ld = [=] __device__(uint64_t k) { return Amp2{amps[insertZeroBit(k, qb)], buf[k]}; };
    auto st = [=] __device__(uint64_t k, const Amp2& v) { amps[insertZeroBit(k, qb)] = make_double2(fma(prob, v.a1.x, v.a0.x), fma(prob, v.a1.y, v.a0.y)); };
    return launchStream<1, Amp2>(half, ld, st);
}

// ---------------------------------------------------------------------------------------------------------
// K17: local_densitymatrix.hpp:134-164. out[l] = sum_{k ascending} in[base(l) | spread(k)]; pure additions in the
// reference's order, so the result is bit-exact. Reads 16*A/2^t, writes 16*A/4^t bytes.
// One thread per output amplitude; neighbouring threads own neighbouring outputs, so for every k a warp reads runs of
// 2^(lowest traced bit) contiguous amplitudes. The 2^t terms of an output are LOADED eight at a time before they are added
// (round 1 issued one dependent load per iteration: 0.30-0.35 of the roofline, latency-bound), and ADDED one by one in
// ascending k, which keeps the sum bit-identical to the reference's loop (:152-158). spreadTab[k] = the k-th term's offset.
template <int BATCH>
__global__ void __launch_bounds__(256) partialTraceKernel(const double2* __restrict__ in, double2* __restrict__ out, uint64_t numOut,
                                                         BitSpec allSorted, const uint64_t* __restrict__ spreadTab, unsigned numTerms) {
    for (uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; l < numOut; l += (uint64_t)gridDim.x * blockDim.x) {
        const double2* base = in + insertZeroBits(l, allSorted);
        double2 acc = make_double2(0.0, 0.0);
        for (unsigned k0 = 0; k0 < numTerms; k0 += BATCH) {
            double2 v[BATCH];
#pragma unroll
            for (int u = 0; u < BATCH; u++) v[u] = base[__ldg(&spreadTab[k0 + u])];
#pragma unroll
            for (int u = 0; u < BATCH; u++) acc = cadd(acc, v[u]);
        }
        out[l] = acc;
    }
}

extern "C" int dfsa_k_partialTrace(dfsa_state* in, dfsa_state* out, const uint32_t* targets, const uint32_t* pairTargets, unsigned numTargets) {
    DFSA_TRY(dfsaEnsureDevice());
    DFSA_REQUIRE(in && out && in->isDensity && out->isDensity && targets && pairTargets, "bad argument");
    DFSA_REQUIRE(numTargets >= 1 && out->numQubits + numTargets == in->numQubits, "output state has the wrong size");
    DFSA_REQUIRE(out->logNumAmps + 2 * numTargets == in->logNumAmps, "output shard has the wrong size");
    BitSpec all, tg, pr;
    std::vector<uint32_t> both(targets, targets + numTargets);
    both.insert(both.end(), pairTargets, pairTargets + numTargets);
    std::sort(both.begin(), both.end());
    all.n = 2 * numTargets; tg.n = pr.n = numTargets;
    for (unsigned q = 0; q < 2 * numTargets; q++) {
        DFSA_REQUIRE(both[q] < in->logNumAmps && (q == 0 || both[q] != both[q - 1]), "traced bits must be distinct suffix bits");
        all.pos[q] = (uint8_t)both[q];
    }
    for (unsigned q = 0; q < numTargets; q++) { tg.pos[q] = (uint8_t)targets[q]; pr.pos[q] = (uint8_t)pairTargets[q]; }
    // offsets of the 2^t terms of one output: bit b of k goes to the ket bit and to the bra bit of target b (:155-156)
    const unsigned numTerms = 1u << numTargets;
    DFSA_REQUIRE(numTargets <= 13, "partialTrace of more than 13 qubits at once");
    std::vector<uint64_t> spread(numTerms);
    for (unsigned k = 0; k < numTerms; k++) {
        uint64_t off = 0;
        for (unsigned b = 0; b < numTargets; b++) if ((k >> b) & 1u) off |= (1ULL << tg.pos[b]) | (1ULL << pr.pos[b]);
        spread[k] = off;
    }
    DfsaContext& c = dfsaCtx();
    void* stage; int slot;
    DFSA_TRY(dfsaStagingAcquire(numTerms * sizeof(uint64_t), &stage, &slot));
    memcpy(stage, spread.data(), numTerms * sizeof(uint64_t));
    double2* scratch;
    DFSA_TRY(dfsaScratch(numTerms * sizeof(uint64_t), &scratch));
    DFSA_CUDA(cudaMemcpyAsync(scratch, stage, numTerms * sizeof(uint64_t), cudaMemcpyHostToDevice, c.compute));
    DFSA_TRY(dfsaStagingCommit(slot));
    const uint64_t* tab = (const uint64_t*)scratch;
    // enough threads to keep ~8 x 16-byte loads per thread x 2048 loads per SM in flight; outputs are few when t is large
    unsigned grid = dfsaGrid(out->numAmps, 256, 1, 8);
    if (numTerms >= 8) partialTraceKernel<8><<<grid, 256, 0, c.compute>>>(in->arr[DFSA_AMPS], out->arr[DFSA_AMPS], out->numAmps, all, tab, numTerms);
    else if (numTerms == 4) partialTraceKernel<4><<<grid, 256, 0, c.compute>>>(in->arr[DFSA_AMPS], out->arr[DFSA_AMPS], out->numAmps, all, tab, numTerms);
    else partialTraceKernel<2><<<grid, 256, 0, c.compute>>>(in->arr[DFSA_AMPS], out->arr[DFSA_AMPS], out->numAmps, all, tab, numTerms);
    DFSA_LAUNCH_CHECK();
    return DFSA_OK;
}

// ---------------------------------------------------------------------------------------------------------
// K23: distributed_densitymatrix.hpp:322-344 with the closed form of getPauliTensorElem (misc.hpp:16-38):
// with lo = row bits, hi = column bits of the flat index, term P contributes only where lo ^ hi == XYmask(P), with
//   elem = i^{#Y} * (-1)^{popc(~hi & Ymask) + popc(hi & Zmask)}          (SURVEY Appendix B).
// GATHER form: each term touches exactly one amplitude per column -> numTerms * 2^N / P scattered 16-byte reads
// (32-byte sectors) instead of a 16*A-byte scan; used when that is the smaller traffic. SCAN form otherwise.
struct PauliTerm { uint64_t xy, y, z; double2 coeff; /* coeff * i^{#Y} */ };

// Block partial -> partials[blockIdx.x]; the LAST block to finish (device-wide ticket) sums the partials in a fixed order
// (deterministic for a given grid) and publishes this rank's value straight into host-visible memory -- the pinned page of
// this process, or its slot of the job's shared page -- followed by a sequence flag. No second kernel, no device-to-host copy,
// no stream synchronisation: the host just watches the flag (three launches + blocking sync + two host barriers).
struct ExpecOut { double* value; unsigned long long* flag; unsigned long long seq; unsigned* ticket; };

__device__ __forceinline__ void blockSum(double& re, double& im) {
    __shared__ double sre[8], sim[8];
    for (int off = 16; off > 0; off >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, off); im += __shfl_xor_sync(0xffffffffu, im, off); }
    __syncthreads();                                     // the scratch may still be read from a previous call in this kernel
    if ((threadIdx.x & 31) == 0) { sre[threadIdx.x >> 5] = re; sim[threadIdx.x >> 5] = im; }
    __syncthreads();
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; w++) { a += sre[w]; b += sim[w]; }
    re = a; im = b;
}

__device__ __forceinline__ void finishExpec(double re, double im, double2* partials, ExpecOut out) {
    blockSum(re, im);
    __shared__ bool last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = make_double2(re, im);
        __threadfence();
        last = (atomicAdd(out.ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double tr = 0.0, ti = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) { const double2 p = __ldcg(&partials[i]); tr += p.x; ti += p.y; }
    blockSum(tr, ti);
    if (threadIdx.x == 0) {
        *out.ticket = 0;                                 // ready for the next call
        volatile double* v = out.value;
        v[0] = tr; v[1] = ti;
        __threadfence_system();
        *(volatile unsigned long long*)out.flag = out.seq;
    }
}

// Item = (column c, a batch of EXPEC_BATCH terms): the batch's amplitudes are LOADED before any of them is used (eight
// independent scattered 16-byte reads in flight per thread), and neighbouring threads take neighbouring batches of the SAME
// column, so a warp's reads stay inside one 16 * 2^N-byte column (same DRAM pages / TLB entries) instead of striding whole
// columns apart. The term table is padded with zero-coefficient terms to a multiple of the batch.
constexpr unsigned EXPEC_BATCH = 8;

__global__ void __launch_bounds__(256) expecGatherKernel(const double2* __restrict__ amps, const PauliTerm* __restrict__ terms, unsigned numBatches,
                                                        unsigned N, unsigned logCols, uint64_t firstCol, double2* partials, ExpecOut out) {
    double re = 0.0, im = 0.0;
    const uint64_t items = (uint64_t)numBatches << logCols;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < items; w += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t c = w / numBatches;
        const PauliTerm* tb = terms + (w - c * numBatches) * EXPEC_BATCH;
        const uint64_t hi = firstCol + c;
        const double2* col = amps + (c << N);
        double2 a[EXPEC_BATCH];
#pragma unroll
        for (unsigned u = 0; u < EXPEC_BATCH; u++) a[u] = col[hi ^ __ldg(&tb[u].xy)];
#pragma unroll
        for (unsigned u = 0; u < EXPEC_BATCH; u++) {
            const PauliTerm tm = tb[u];
            const unsigned neg = (unsigned)(__popcll(~hi & tm.y) + __popcll(hi & tm.z)) & 1u;
            const double2 v = cmul(tm.coeff, a[u]);
            re += neg ? -v.x : v.x;
            im += neg ? -v.y : v.y;
        }
    }
    finishExpec(re, im, partials, out);
}

__global__ void __launch_bounds__(256) expecScanKernel(const double2* __restrict__ amps, const PauliTerm* __restrict__ terms, unsigned numTerms,
                                                      unsigned N, uint64_t numAmps, uint64_t rankShift, double2* partials, ExpecOut out) {
    double re = 0.0, im = 0.0;
    const uint64_t loMask = (1ULL << N) - 1ULL;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < numAmps; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = rankShift | j, lo = i & loMask, hi = i >> N, x = lo ^ hi;
        double2 sum = make_double2(0.0, 0.0);
        for (unsigned t = 0; t < numTerms; t++) {
            const PauliTerm tm = terms[t];
            if (tm.xy == x) {
                const unsigned neg = (unsigned)(__popcll(~hi & tm.y) + __popcll(hi & tm.z)) & 1u;
                sum.x += neg ? -tm.coeff.x : tm.coeff.x;
                sum.y += neg ? -tm.coeff.y : tm.coeff.y;
            }
        }
        if (sum.x != 0.0 || sum.y != 0.0) {
            double2 v = cmul(sum, amps[j]);
            re += v.x; im += v.y;
        }
    }
    finishExpec(re, im, partials, out);
}

// K23 + X11: the expectation value summed over all ranks (distributed_densitymatrix.hpp:322-344 incl. comm_reduceAmp).
// *outIsGlobal = 1: out[] is the global sum already (every rank read every rank's slot of the shared page, rank order);
// 0: out[] is this rank's part, combine with dfsa_x_allreduce_amp.
extern "C" int dfsa_kx_expecPauliString(dfsa_state* s, const double* coeffs, unsigned numTerms, const uint32_t* paulis, double out[2], int* outIsGlobal) {
    DFSA_DM_ENTRY(s);
    DFSA_REQUIRE(coeffs && paulis && out && numTerms >= 1, "bad argument");
    const unsigned N = s->numQubits;
    const unsigned numBatches = (numTerms + EXPEC_BATCH - 1) / EXPEC_BATCH, numPadded = numBatches * EXPEC_BATCH;
    std::vector<PauliTerm> terms(numPadded, PauliTerm{0, 0, 0, make_double2(0.0, 0.0)});     // padding: coefficient 0
    for (unsigned t = 0; t < numTerms; t++) {
        PauliTerm tm{0, 0, 0, make_double2(coeffs[t], 0.0)};
        unsigned numY = 0;
        for (unsigned q = 0; q < N; q++) {
            uint32_t code = paulis[(size_t)t * N + q];
            DFSA_REQUIRE(code <= 3, "Pauli codes are 0..3 (types.hpp:16-18)");
            if (code == 1 || code == 2) tm.xy |= 1ULL << q;
            if (code == 2) { tm.y |= 1ULL << q; numY++; }
            if (code == 3) tm.z |= 1ULL << q;
        }
        switch (numY & 3u) { case 1: tm.coeff = make_double2(0.0, coeffs[t]); break; case 2: tm.coeff = make_double2(-coeffs[t], 0.0); break;
                             case 3: tm.coeff = make_double2(0.0, -coeffs[t]); break; default: break; }
        terms[t] = tm;
    }
    DfsaContext& c = dfsaCtx();
    const unsigned logCols = s->logNumAmps - N;                 // columns owned by this rank
    bool gather = ((uint64_t)numTerms << 1) <= (1ULL << N);     // 32-byte sectors per gathered amp vs a 16*A scan
    if (getenv("DFSA_EXPEC_FORCE_GATHER")) gather = true;
    if (getenv("DFSA_EXPEC_FORCE_SCAN")) gather = false;
    const bool scan = !gather;
    uint64_t items = scan ? s->numAmps : ((uint64_t)numBatches << logCols);
    unsigned grid = dfsaGrid(items, 256, 1, 8);
    size_t termBytes = (sizeof(PauliTerm) * numPadded + 255) / 256 * 256;
    double2* scratch;
    DFSA_TRY(dfsaScratch(termBytes + (grid + 1) * sizeof(double2), &scratch));
    PauliTerm* dTerms = (PauliTerm*)scratch;
    double2* partials = (double2*)((char*)scratch + termBytes);
    // term table: through the pinned staging ring when it fits (asynchronous), else a blocking pageable copy
    void* stage = nullptr; int slot = -1;
    if (dfsaStagingAcquire(sizeof(PauliTerm) * numPadded, &stage, &slot) == DFSA_OK) {
        memcpy(stage, terms.data(), sizeof(PauliTerm) * numPadded);
        DFSA_CUDA(cudaMemcpyAsync(dTerms, stage, sizeof(PauliTerm) * numPadded, cudaMemcpyHostToDevice, c.compute));
        DFSA_TRY(dfsaStagingCommit(slot));
    } else {
        DFSA_CUDA(cudaMemcpyAsync(dTerms, terms.data(), sizeof(PauliTerm) * numPadded, cudaMemcpyHostToDevice, c.compute));
        DFSA_CUDA(cudaStreamSynchronize(c.compute));
    }
    // where the last block publishes: this rank's slot of the job's shared page when the device can write there (the sum over
    // ranks is then formed by every host reading all slots -- out[] is already the GLOBAL value and dfsa_x_allreduce_amp must
    // not be applied again: *outIsGlobal), else this process's pinned page
    ExpecOut eo;
    int global = 0;
    DFSA_TRY(dfsaExpecTarget(&eo.value, &eo.flag, &eo.seq, &eo.ticket, &global));
    if (scan) expecScanKernel<<<grid, 256, 0, c.compute>>>(s->arr[DFSA_AMPS], dTerms, numTerms, N, s->numAmps, (uint64_t)s->rank << s->logNumAmps, partials, eo);
    else expecGatherKernel<<<grid, 256, 0, c.compute>>>(s->arr[DFSA_AMPS], dTerms, numBatches, N, logCols, (uint64_t)s->rank << logCols, partials, eo);
    DFSA_LAUNCH_CHECK();
    DFSA_TRY(dfsaExpecCollect(eo.seq, out));
    if (outIsGlobal) *outIsGlobal = global;
    else if (global) { dfsaSetError("internal: global expectation value needs the outIsGlobal form"); return DFSA_ERR_ARG; }
    return DFSA_OK;
}

extern "C" int dfsa_k_expecPauliString(dfsa_state* s, const double* coeffs, unsigned numTerms, const uint32_t* paulis, double out[2]) {
    // local part only (the documented contract of this entry): keep the cross-rank sum out of it
    int global = 0;
    DFSA_TRY(dfsaExpecLocalOnly(true));
    int rc = dfsa_kx_expecPauliString(s, coeffs, numTerms, paulis, out, &global);
    dfsaExpecLocalOnly(false);
    return rc;
}
