// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the arithmetic the reference
// (valida-xyz/valida @ 5058de85) reaches through its un-vendored dependency
// valida-xyz/Plonky3 @ bdd338d61b3c1f24f17abd3b2848f1c910c49428 (Cargo.toml:24-41, Cargo.lock:651-870).
// PARITY UNPINNED: the Plonky3 source is absent from /root/reference, and the reference holds no
// golden vectors for this path; the published algorithm is restated here.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// use anything under oracle/.  The product path (valida_b200/) never links or calls it.
//
// BabyBear: p = 2^31 - 2^27 + 1 (p3-baby-bear), values held in CANONICAL form here (the GPU side
// works in Montgomery form: a different representation on purpose).
// Ext5: BinomialExtensionField<BabyBear,5>, X^5 = W = 2 (basic/src/bin/valida.rs:357-358).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include <array>
#include <cassert>
#include <algorithm>

namespace orc {

constexpr uint32_t P = 2013265921u;       // 0x78000001
constexpr uint32_t GEN = 31;              // BabyBear::generator()
constexpr uint32_t TWO_ADIC_ROOT_27 = 0x1a427a41u;  // 31^15: generator of the order-2^27 subgroup
constexpr int TWO_ADICITY = 27;

static inline uint32_t add(uint32_t a, uint32_t b) { uint32_t s = a + b; return s >= P ? s - P : s; }
static inline uint32_t sub(uint32_t a, uint32_t b) { return a >= b ? a - b : a + P - b; }
static inline uint32_t neg(uint32_t a) { return a ? P - a : 0; }
static inline uint32_t mul(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) % P); }
static inline uint32_t from_u64(uint64_t x) { return (uint32_t)(x % P); }
static inline uint32_t from_i64(int64_t x) { int64_t r = x % (int64_t)P; if (r < 0) r += P; return (uint32_t)r; }
static inline uint32_t pw(uint32_t a, uint64_t e) {
    uint32_t r = 1;
    while (e) { if (e & 1) r = mul(r, a); a = mul(a, a); e >>= 1; }
    return r;
}
static inline uint32_t inv(uint32_t a) { assert(a != 0); return pw(a, P - 2); }
static inline uint32_t exp_pow2(uint32_t a, int k) { while (k-- > 0) a = mul(a, a); return a; }
// TwoAdicField::two_adic_generator(bits)  [P3-UNVERIFIED; SURVEY App. A item 2]
static inline uint32_t two_adic_generator(int bits) { assert(bits <= TWO_ADICITY); return exp_pow2(TWO_ADIC_ROOT_27, TWO_ADICITY - bits); }

// out[i] = start * base^i for i < n, on all host threads (each chunk restarts from start * base^chunk_begin): replaces the serial
// running products that made up most of the prover's non-parallel time.
static inline std::vector<uint32_t> geometric(uint32_t start, uint32_t base, size_t n) {
    std::vector<uint32_t> out(n);
    const size_t CH = 1 << 14;
    const long nch = (long)((n + CH - 1) / CH);
#pragma omp parallel for schedule(static) if (n > 4 * CH)
    for (long c = 0; c < nch; c++) {
        const size_t a = (size_t)c * CH, b = std::min(n, a + CH);
        uint32_t x = mul(start, pw(base, a));
        for (size_t i = a; i < b; i++) { out[i] = x; x = mul(x, base); }
    }
    return out;
}

static inline uint32_t reverse_bits_len(uint32_t x, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | ((x >> i) & 1); }
    return r;
}
static inline int log2_strict(uint64_t n) { int l = 0; while ((1ull << l) < n) l++; assert((1ull << l) == n); return l; }
static inline int log2_ceil(uint64_t n) { int l = 0; while ((1ull << l) < n) l++; return l; }

// ---------------------------------------------------------------------------------------------
struct Ext5 {
    uint32_t c[5];
    static Ext5 zero() { Ext5 r; std::memset(r.c, 0, sizeof r.c); return r; }
    static Ext5 one() { Ext5 r = zero(); r.c[0] = 1; return r; }
    static Ext5 from_base(uint32_t b) { Ext5 r = zero(); r.c[0] = b; return r; }
    static Ext5 monomial(int i) { Ext5 r = zero(); r.c[i] = 1; return r; }
    bool is_zero() const { return !(c[0] | c[1] | c[2] | c[3] | c[4]); }
    bool operator==(const Ext5& o) const { return std::memcmp(c, o.c, sizeof c) == 0; }
    bool operator!=(const Ext5& o) const { return !(*this == o); }
};
constexpr uint32_t EXT_W = 2;

static inline Ext5 operator+(const Ext5& a, const Ext5& b) { Ext5 r; for (int i = 0; i < 5; i++) r.c[i] = add(a.c[i], b.c[i]); return r; }
static inline Ext5 operator-(const Ext5& a, const Ext5& b) { Ext5 r; for (int i = 0; i < 5; i++) r.c[i] = sub(a.c[i], b.c[i]); return r; }
static inline Ext5 operator-(const Ext5& a) { Ext5 r; for (int i = 0; i < 5; i++) r.c[i] = neg(a.c[i]); return r; }
static inline Ext5 operator*(const Ext5& a, const Ext5& b) {
    // schoolbook product reduced with X^5 = 2.  Canonical factors are < p < 2^31, so a raw product is < 2^62 and FOUR of them
    // fit in a u64 (4 p^2 < 2^64): every diagonal is summed raw, at most four terms at a time, and reduced once or twice —
    // 11 reductions instead of one per term (same values: reduction is exact).
    const uint32_t* x = a.c; const uint32_t* y = b.c;
    auto pr = [&](int i, int j) { return (uint64_t)x[i] * y[j]; };
    const uint64_t t0 = pr(0, 0) % P;
    const uint64_t t1 = (pr(0, 1) + pr(1, 0)) % P;
    const uint64_t t2 = (pr(0, 2) + pr(1, 1) + pr(2, 0)) % P;
    const uint64_t t3 = (pr(0, 3) + pr(1, 2) + pr(2, 1) + pr(3, 0)) % P;
    const uint64_t t4 = ((pr(0, 4) + pr(1, 3) + pr(2, 2) + pr(3, 1)) % P + pr(4, 0)) % P;
    const uint64_t t5 = (pr(1, 4) + pr(2, 3) + pr(3, 2) + pr(4, 1)) % P;
    const uint64_t t6 = (pr(2, 4) + pr(3, 3) + pr(4, 2)) % P;
    const uint64_t t7 = (pr(3, 4) + pr(4, 3)) % P;
    const uint64_t t8 = pr(4, 4) % P;
    Ext5 r;
    r.c[0] = (uint32_t)((t0 + EXT_W * t5) % P);
    r.c[1] = (uint32_t)((t1 + EXT_W * t6) % P);
    r.c[2] = (uint32_t)((t2 + EXT_W * t7) % P);
    r.c[3] = (uint32_t)((t3 + EXT_W * t8) % P);
    r.c[4] = (uint32_t)t4;
    return r;
}
static inline Ext5 operator*(const Ext5& a, uint32_t s) { Ext5 r; for (int i = 0; i < 5; i++) r.c[i] = mul(a.c[i], s); return r; }
static inline Ext5& operator+=(Ext5& a, const Ext5& b) { a = a + b; return a; }
static inline Ext5& operator-=(Ext5& a, const Ext5& b) { a = a - b; return a; }
static inline Ext5& operator*=(Ext5& a, const Ext5& b) { a = a * b; return a; }
static inline Ext5 operator+(const Ext5& a, uint32_t b) { Ext5 r = a; r.c[0] = add(r.c[0], b); return r; }
static inline Ext5 operator-(const Ext5& a, uint32_t b) { Ext5 r = a; r.c[0] = sub(r.c[0], b); return r; }

// Sum of ext5 x base products with the reductions postponed: four raw products (each < p^2 < 2^61.9) plus a canonical
// remainder fit in a u64, so fold() is due after every fourth mad().  Same values as the reduced sum (reduction is exact).
struct Lazy5 {
    uint64_t s[5] = {0, 0, 0, 0, 0};
    void mad(const Ext5& a, uint32_t b) { for (int i = 0; i < 5; i++) s[i] += (uint64_t)a.c[i] * b; }
    void fold() { for (int i = 0; i < 5; i++) s[i] %= P; }
    Ext5 value() const { Ext5 r; for (int i = 0; i < 5; i++) r.c[i] = (uint32_t)(s[i] % P); return r; }
};

static inline Ext5 ext_pow(Ext5 a, uint64_t e) {
    Ext5 r = Ext5::one();
    while (e) { if (e & 1) r = r * a; a = a * a; e >>= 1; }
    return r;
}
static inline Ext5 ext_exp_pow2(Ext5 a, int k) { while (k-- > 0) a = a * a; return a; }

// Frobenius x -> x^p on F_p[X]/(X^5-2): coefficient i is scaled by z^i, z = W^((p-1)/5).
static inline Ext5 frobenius(const Ext5& a) {
    static const uint32_t z = pw(EXT_W, (P - 1) / 5);
    Ext5 r; uint32_t zi = 1;
    for (int i = 0; i < 5; i++) { r.c[i] = mul(a.c[i], zi); zi = mul(zi, z); }
    return r;
}
// a^{-1} = f1 f2 f3 f4 / Norm(a),  f_k = Frobenius^k(a), Norm(a) = a f1 f2 f3 f4 in F_p.
static inline Ext5 ext_inv(const Ext5& a) {
    assert(!a.is_zero());
    Ext5 f1 = frobenius(a), f2 = frobenius(f1), f3 = frobenius(f2), f4 = frobenius(f3);
    Ext5 prod = f1 * f2 * f3 * f4;
    Ext5 n = a * prod;
    assert(n.c[1] == 0 && n.c[2] == 0 && n.c[3] == 0 && n.c[4] == 0);
    return prod * inv(n.c[0]);
}
static inline Ext5 operator/(const Ext5& a, const Ext5& b) { return a * ext_inv(b); }

// p3_field::batch_multiplicative_inverse (no zeros allowed)
template <class T, class Inv, class One>
static inline std::vector<T> batch_inverse_generic(const std::vector<T>& v, Inv invf, One one) {
    size_t n = v.size();
    std::vector<T> out(n);
    if (!n) return out;
    std::vector<T> pref(n);
    T acc = one;
    for (size_t i = 0; i < n; i++) { pref[i] = acc; acc = acc * v[i]; }
    T ia = invf(acc);
    for (size_t i = n; i-- > 0;) { out[i] = ia * pref[i]; ia = ia * v[i]; }
    return out;
}
struct MulU32 { uint32_t v; };
static inline std::vector<uint32_t> batch_inverse(const std::vector<uint32_t>& v) {
    size_t n = v.size();
    std::vector<uint32_t> out(n);
    const size_t CH = 8192;
    long nch = (long)((n + CH - 1) / CH);
#pragma omp parallel for schedule(static) if (n > 4 * CH)
    for (long c = 0; c < nch; c++) {
        size_t a = (size_t)c * CH, b = std::min(n, a + CH);
        std::vector<uint32_t> pref(b - a);
        uint32_t acc = 1;
        for (size_t i = a; i < b; i++) { pref[i - a] = acc; acc = mul(acc, v[i]); }
        uint32_t ia = inv(acc);
        for (size_t i = b; i-- > a;) { out[i] = mul(ia, pref[i - a]); ia = mul(ia, v[i]); }
    }
    return out;
}
// chunked so that the Montgomery trick runs on all cores (the reference's p3_field version is serial;
// results are identical — inversion is exact)
static inline std::vector<Ext5> batch_inverse(const std::vector<Ext5>& v) {
    size_t n = v.size();
    std::vector<Ext5> out(n);
    const size_t CH = 4096;
    long nch = (long)((n + CH - 1) / CH);
#pragma omp parallel for schedule(static) if (n > 4 * CH)
    for (long c = 0; c < nch; c++) {
        size_t a = (size_t)c * CH, b = std::min(n, a + CH);
        std::vector<Ext5> pref(b - a);
        Ext5 acc = Ext5::one();
        for (size_t i = a; i < b; i++) { pref[i - a] = acc; acc = acc * v[i]; }
        Ext5 ia = ext_inv(acc);
        for (size_t i = b; i-- > a;) { out[i] = ia * pref[i - a]; ia = ia * v[i]; }
    }
    return out;
}
// valida util::batch_multiplicative_inverse_allowing_zero (util/src/lib.rs:21-43): zeros stay zero.
// Chunked like batch_inverse, zeros skipped in place (no serial gather / scatter of the non-zero entries).
static inline std::vector<Ext5> batch_inverse_allowing_zero(const std::vector<Ext5>& v) {
    size_t n = v.size();
    std::vector<Ext5> out(n);
    const size_t CH = 4096;
    long nch = (long)((n + CH - 1) / CH);
#pragma omp parallel for schedule(static) if (n > 4 * CH)
    for (long c = 0; c < nch; c++) {
        size_t a = (size_t)c * CH, b = std::min(n, a + CH);
        std::vector<Ext5> pref(b - a);
        Ext5 acc = Ext5::one();
        bool any = false;
        for (size_t i = a; i < b; i++) { pref[i - a] = acc; if (!v[i].is_zero()) { acc = acc * v[i]; any = true; } }
        Ext5 ia = any ? ext_inv(acc) : Ext5::one();
        for (size_t i = b; i-- > a;) {
            if (v[i].is_zero()) { out[i] = Ext5::zero(); continue; }
            out[i] = ia * pref[i - a]; ia = ia * v[i];
        }
    }
    return out;
}
static inline std::vector<uint32_t> batch_inverse_allowing_zero(const std::vector<uint32_t>& v) {
    size_t n = v.size();
    std::vector<uint32_t> out(n);
    const size_t CH = 8192;
    long nch = (long)((n + CH - 1) / CH);
#pragma omp parallel for schedule(static) if (n > 4 * CH)
    for (long c = 0; c < nch; c++) {
        size_t a = (size_t)c * CH, b = std::min(n, a + CH);
        std::vector<uint32_t> pref(b - a);
        uint32_t acc = 1;
        for (size_t i = a; i < b; i++) { pref[i - a] = acc; if (v[i]) acc = mul(acc, v[i]); }
        uint32_t ia = inv(acc);
        for (size_t i = b; i-- > a;) {
            if (!v[i]) { out[i] = 0; continue; }
            out[i] = mul(ia, pref[i - a]); ia = mul(ia, v[i]);
        }
    }
    return out;
}

// Row-major matrix of base-field elements (p3_matrix::dense::RowMajorMatrix<Val>).
struct Matrix {
    std::vector<uint32_t> v;
    size_t width = 0;
    Matrix() {}
    Matrix(size_t h, size_t w) : v(h * w, 0), width(w) {}
    Matrix(std::vector<uint32_t> vals, size_t w) : v(std::move(vals)), width(w) {}
    size_t height() const { return width ? v.size() / width : 0; }
    uint32_t* row(size_t r) { return v.data() + r * width; }
    const uint32_t* row(size_t r) const { return v.data() + r * width; }
    uint32_t& at(size_t r, size_t c) { return v[r * width + c]; }
    uint32_t at(size_t r, size_t c) const { return v[r * width + c]; }
};

}  // namespace orc
