// types.hpp -- precision typedefs and small dense-matrix helpers of the drop-in host API.
// Same public names as the reference's src/types.hpp (:16-18 PauliOperator, :26-29 Real/Nat/Index,
// :37-42 Amp/AmpArray/AmpMatrix/NatArray/MatrixArray/RealArray, :58-149 helper algebra) so user code compiles
// unchanged; the bodies are this project's own. Everything here is host-only and tiny (gate matrices, planning).
#pragma once

#include <cassert>
#include <complex>
#include <cstddef>
#include <vector>

enum PauliOperator { I = 0, X = 1, Y = 2, Z = 3 };

using Real  = double;
using Nat   = unsigned int;
using Index = unsigned long long;
using Amp   = std::complex<Real>;

using AmpArray    = std::vector<Amp>;
using AmpMatrix   = std::vector<AmpArray>;
using NatArray    = std::vector<Nat>;
using MatrixArray = std::vector<AmpMatrix>;
using RealArray   = std::vector<Real>;

inline AmpMatrix getZeroMatrix(Index dim) {
    assert(dim > 0);
    return AmpMatrix(dim, AmpArray(dim, Amp(0, 0)));
}

inline AmpMatrix getIdentityMatrix(Index dim) {
    AmpMatrix m = getZeroMatrix(dim);
    for (Index d = 0; d < dim; d++) m[d][d] = Amp(1, 0);
    return m;
}

inline AmpMatrix getConjugateMatrix(const AmpMatrix& in) {
    AmpMatrix out = in;
    for (AmpArray& row : out)
        for (Amp& e : row) e = std::conj(e);
    return out;
}

inline AmpMatrix getDaggerMatrix(const AmpMatrix& in) {
    const std::size_t n = in.size();
    AmpMatrix out = getZeroMatrix(n);
    for (std::size_t r = 0; r < n; r++)
        for (std::size_t c = 0; c < n; c++) out[c][r] = std::conj(in[r][c]);
    return out;
}

inline AmpMatrix operator*(const Amp& scalar, const AmpMatrix& m) {
    AmpMatrix out = m;
    for (AmpArray& row : out)
        for (Amp& e : row) e *= scalar;
    return out;
}

inline AmpMatrix operator*(const AmpMatrix& a, const AmpMatrix& b) {
    assert(a.size() == b.size());
    const std::size_t n = a.size();
    AmpMatrix out = getZeroMatrix(n);
    for (std::size_t r = 0; r < n; r++)
        for (std::size_t k = 0; k < n; k++) {
            const Amp ark = a[r][k];
            for (std::size_t c = 0; c < n; c++) out[r][c] += ark * b[k][c];
        }
    return out;
}

inline AmpMatrix operator+(const AmpMatrix& a, const AmpMatrix& b) {
    assert(a.size() == b.size());
    AmpMatrix out = a;
    for (std::size_t r = 0; r < out.size(); r++)
        for (std::size_t c = 0; c < out.size(); c++) out[r][c] += b[r][c];
    return out;
}

// Kronecker product a (x) b: b indexes the low bits
inline AmpMatrix operator%(const AmpMatrix& a, const AmpMatrix& b) {
    const std::size_t na = a.size(), nb = b.size();
    AmpMatrix out = getZeroMatrix(na * nb);
    for (std::size_t ra = 0; ra < na; ra++)
        for (std::size_t ca = 0; ca < na; ca++)
            for (std::size_t rb = 0; rb < nb; rb++)
                for (std::size_t cb = 0; cb < nb; cb++) out[ra * nb + rb][ca * nb + cb] = a[ra][ca] * b[rb][cb];
    return out;
}

inline AmpArray operator*(const AmpMatrix& m, const AmpArray& v) {
    assert(m.size() == v.size());
    AmpArray out(v.size(), Amp(0, 0));
    for (std::size_t r = 0; r < v.size(); r++)
        for (std::size_t c = 0; c < v.size(); c++) out[r] += m[r][c] * v[c];
    return out;
}

// row-major interleaved (re,im) doubles, the layout the C-ABI takes
inline std::vector<double> dfsaFlatten(const AmpMatrix& m) {
    std::vector<double> flat;
    flat.reserve(2 * m.size() * m.size());
    for (const AmpArray& row : m) {
        assert(row.size() == m.size());
        for (const Amp& e : row) { flat.push_back(e.real()); flat.push_back(e.imag()); }
    }
    return flat;
}
