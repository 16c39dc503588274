// bit_maths.hpp -- host-side index twiddles of the drop-in API (reference: src/bit_maths.hpp:19-173).
// 64-bit throughout (the reference's Nat shifts truncate above 32 index bits) and with setBit
// clearing the addressed bit before writing it (the reference's :66 keeps only that bit).
// The device kernels carry their own copies (csrc/dfsa_internal.cuh); these serve host planning and user code.
#pragma once

#include "types.hpp"

inline Index powerOf2(Nat exponent) { return Index(1) << exponent; }
inline bool  isPowerOf2(Index number) { return number != 0 && (number & (number - 1)) == 0; }
inline Nat   getBit(Index number, Nat bitIndex) { return Nat((number >> bitIndex) & Index(1)); }
inline Index flipBit(Index number, Nat bitIndex) { return number ^ (Index(1) << bitIndex); }

inline Index insertBit(Index number, Nat bitIndex, Nat bitValue) {
    const Index below = number & ((Index(1) << bitIndex) - 1);
    return ((number ^ below) << 1) | (Index(bitValue & 1u) << bitIndex) | below;
}

// bitIndices strictly increasing
inline Index insertBits(Index number, const NatArray& bitIndices, Nat bitValue) {
    for (Nat pos : bitIndices) number = insertBit(number, pos, bitValue);
    return number;
}

inline Index setBit(Index number, Nat bitIndex, Nat bitValue) {
    const Index bit = Index(1) << bitIndex;
    return (number & ~bit) | (Index(bitValue & 1u) << bitIndex);
}

// bit q of bitsValue goes to position bitIndices[q]
inline Index setBits(Index number, const NatArray& bitIndices, Index bitsValue) {
    for (std::size_t q = 0; q < bitIndices.size(); q++) number = setBit(number, bitIndices[q], getBit(bitsValue, Nat(q)));
    return number;
}

inline Nat getBitMaskParity(Index mask) { return Nat(__builtin_parityll(mask)); }

inline Index insertTwoBits(Index number, Nat highInd, Nat highBit, Nat lowInd, Nat lowBit) {
    return insertBit(insertBit(number, lowInd, lowBit), highInd, highBit);
}
inline Index insertThreeZeroBits(Index number, Nat i3, Nat i2, Nat i1) { return insertBit(insertTwoBits(number, i2, 0, i1, 0), i3, 0); }
inline Index insertFourZeroBits(Index number, Nat i4, Nat i3, Nat i2, Nat i1) { return insertTwoBits(insertTwoBits(number, i2, 0, i1, 0), i4, 0, i3, 0); }
inline Index flipTwoBits(Index number, Nat i1, Nat i0) { return number ^ (Index(1) << i1) ^ (Index(1) << i0); }

// first zero bit of mask strictly below bitInd, scanning downwards
inline Nat getNextLeftmostZeroBit(Index mask, Nat bitInd) {
    do { bitInd--; } while (getBit(mask, bitInd));
    return bitInd;
}

inline bool allBitsAreOne(Index number, const NatArray& bitIndices) {
    for (Nat pos : bitIndices)
        if (!getBit(number, pos)) return false;
    return true;
}

inline Index getBitMask(const NatArray& bitIndices) {
    Index mask = 0;
    for (Nat pos : bitIndices) mask ^= Index(1) << pos;
    return mask;
}

inline Nat logBase2(Index powerOf2Value) {
    Nat e = 0;
    while (!(powerOf2Value & 1)) { powerOf2Value >>= 1; e++; }
    return e;
}
