// ORACLE — TEST INFRASTRUCTURE ONLY (see field.h header).
// TwoAdicSubgroupDft::{dft_batch, idft_batch, coset_lde_batch} of p3-dft, as reached from
// TwoAdicFriPcs::commit_shifted_batches (reference call sites derive/src/lib.rs:309,330,355,372;
// DFT type chosen at basic/src/bin/valida.rs:379 / basic/tests/test_prover.rs:436).
// Field arithmetic is exact, so Radix2DitParallel and Radix2Bowers give identical outputs;
// this restates the mathematical transform: out[k] = sum_j in[j] * w^(j k), w = two_adic_generator(log n),
// natural order in and out, per column of a row-major matrix.  [P3-UNVERIFIED; App. A item 4]
#pragma once
#include "field.h"
#include <memory>

namespace orc {

static inline void bit_reverse_rows(Matrix& m) {
    size_t h = m.height(), w = m.width;
    if (h <= 1) return;
    int lg = log2_strict(h);
#pragma omp parallel for schedule(static) if (h * w > (1u << 16))
    for (long i = 0; i < (long)h; i++) {
        size_t j = reverse_bits_len((uint32_t)i, lg);
        if ((size_t)i < j) for (size_t c = 0; c < w; c++) std::swap(m.v[i * w + c], m.v[j * w + c]);
    }
}

// ---- how the transform is computed (the outputs are the ones stated above, bit for bit) ----
// The columns are moved to column-major scratch so that a butterfly touches contiguous words, the first 14 stages of a
// column run block by block (64 KB: one block stays in the core's cache for all of them), only the remaining
// stages sweep the whole column, twiddles are kept in Montgomery form (one reduction per product, data stays
// canonical), and the work is spread over the host threads as (column, block) pairs.  A CPU prover of the
// reference's class (p3-dft's Radix2DitParallel on rayon) does the same things; the first version of this file swept
// the whole row-major matrix once per stage with a 64-bit `%` per product and was 4-5x slower.
namespace nttimpl {

constexpr uint32_t newton_pinv() { uint32_t x = 1; for (int i = 0; i < 5; i++) x *= 2u - P * x; return x; }
static constexpr uint32_t PINV = newton_pinv();                    // P * PINV == 1 (mod 2^32)
static_assert((uint32_t)(P * PINV) == 1u, "Montgomery constant");
static inline uint32_t to_mont(uint32_t a) { return (uint32_t)(((uint64_t)a << 32) % P); }
// a * b / 2^32 mod P, canonical, for a, b < P: with b = to_mont(t) this is a * t mod P.
static inline uint32_t mmul(uint32_t a, uint32_t bm) {
    uint64_t t = (uint64_t)a * bm;
    uint32_t q = (uint32_t)t * PINV;
    uint32_t hi = (uint32_t)(t >> 32), qp = (uint32_t)(((uint64_t)q * P) >> 32);
    uint32_t r = hi - qp;
    return hi < qp ? r + P : r;
}

constexpr int LG_BLOCK = 14;

// Twiddles of one size: stage s (butterfly span 2^s) uses w_s^k = root^(k * 2^(lg - s)), k < 2^(s-1), stored at
// tw[2^(s-1) - 1 + k] in Montgomery form.
struct Plan {
    int lg = 0;
    std::vector<uint32_t> tw;
    const uint32_t* stage(int s) const { return tw.data() + ((size_t(1) << (s - 1)) - 1); }
};
static inline Plan make_plan(int lg, bool inverse) {
    Plan p; p.lg = lg;
    if (lg == 0) return p;
    size_t n = size_t(1) << lg;
    uint32_t root = two_adic_generator(lg);
    if (inverse) root = inv(root);
    p.tw.resize(n - 1);
    uint32_t* top = p.tw.data() + (n / 2 - 1);
    const size_t half = n / 2, CH = 4096;                           // powers of root, chunked so the threads share the work
    long nch = (long)((half + CH - 1) / CH);
#pragma omp parallel for schedule(static) if (half > (1u << 16))
    for (long ch = 0; ch < nch; ch++) {
        size_t k0 = (size_t)ch * CH, k1 = std::min(half, k0 + CH);
        uint32_t x = pw(root, k0);
        for (size_t k = k0; k < k1; k++) { top[k] = to_mont(x); x = mul(x, root); }
    }
    for (int s = lg - 1; s >= 1; s--) {
        const uint32_t* up = p.stage(s + 1);
        uint32_t* me = p.tw.data() + ((size_t(1) << (s - 1)) - 1);
        size_t hs = size_t(1) << (s - 1);
        for (size_t k = 0; k < hs; k++) me[k] = up[2 * k];
    }
    return p;
}

// Decimation in time, stage s on the butterflies [b0, b1) of one column: (x, y) -> (x + t y, x - t y).
static inline void dit_span(uint32_t* a, const uint32_t* tw, size_t half, size_t b0, size_t b1) {
    while (b0 < b1) {
        size_t blk = b0 / half, k0 = b0 % half, cnt = std::min(half - k0, b1 - b0);
        uint32_t* x = a + blk * 2 * half + k0;
        uint32_t* y = x + half;
        const uint32_t* t = tw + k0;
#pragma omp simd
        for (size_t k = 0; k < cnt; k++) {
            uint32_t u = x[k], v = mmul(y[k], t[k]);
            uint32_t s = u + v, d = u - v;
            x[k] = s >= P ? s - P : s;
            y[k] = u < v ? d + P : d;
        }
        b0 += cnt;
    }
}
// Decimation in frequency: (x, y) -> (x + y, (x - y) t).
static inline void dif_span(uint32_t* a, const uint32_t* tw, size_t half, size_t b0, size_t b1) {
    while (b0 < b1) {
        size_t blk = b0 / half, k0 = b0 % half, cnt = std::min(half - k0, b1 - b0);
        uint32_t* x = a + blk * 2 * half + k0;
        uint32_t* y = x + half;
        const uint32_t* t = tw + k0;
#pragma omp simd
        for (size_t k = 0; k < cnt; k++) {
            uint32_t u = x[k], v = y[k];
            uint32_t s = u + v, d = u - v;
            x[k] = s >= P ? s - P : s;
            y[k] = mmul(u < v ? d + P : d, t[k]);
        }
        b0 += cnt;
    }
}

// cols: w columns of n = 2^lg words each.  DIT: bit-reversed order in, natural order out.  DIF: natural in, bit-reversed out.
static inline void transform_columns(uint32_t* cols, size_t w, const Plan& p, bool dif) {
    const int lg = p.lg;
    if (lg == 0 || w == 0) return;
    const size_t n = size_t(1) << lg;
    const int lgb = std::min(lg, LG_BLOCK);
    const size_t B = size_t(1) << lgb, nblk = n / B;
    const long tasks = (long)(w * nblk);
    const bool par = w * n > (1u << 16);
    auto block_stages = [&]() {
#pragma omp parallel for schedule(static) if (par)
        for (long t = 0; t < tasks; t++) {
            uint32_t* a = cols + (size_t)t * B;                     // columns are contiguous: task t = (column, block)
            if (!dif) for (int s = 1; s <= lgb; s++) dit_span(a, p.stage(s), size_t(1) << (s - 1), 0, B / 2);
            else for (int s = lgb; s >= 1; s--) dif_span(a, p.stage(s), size_t(1) << (s - 1), 0, B / 2);
        }
    };
    auto sweep_stage = [&](int s) {                                 // span > block: chunks of B/2 butterflies of one column
#pragma omp parallel for schedule(static) if (par)
        for (long t = 0; t < tasks; t++) {
            size_t c = (size_t)t / nblk, ch = (size_t)t % nblk;
            uint32_t* a = cols + c * n;
            if (!dif) dit_span(a, p.stage(s), size_t(1) << (s - 1), ch * (B / 2), (ch + 1) * (B / 2));
            else dif_span(a, p.stage(s), size_t(1) << (s - 1), ch * (B / 2), (ch + 1) * (B / 2));
        }
    };
    if (!dif) { block_stages(); for (int s = lgb + 1; s <= lg; s++) sweep_stage(s); }
    else { for (int s = lg; s > lgb; s--) sweep_stage(s); block_stages(); }
}

// Row-major h x w  <->  column-major w x h.
static inline void to_columns(const uint32_t* rows, size_t h, size_t w, uint32_t* cols) {
    const size_t T = 512;
    long nt = (long)((h + T - 1) / T);
#pragma omp parallel for schedule(static) if (h * w > (1u << 16))
    for (long t = 0; t < nt; t++) {
        size_t r0 = (size_t)t * T, r1 = std::min(h, r0 + T);
        for (size_t c = 0; c < w; c++) { uint32_t* o = cols + c * h; for (size_t r = r0; r < r1; r++) o[r] = rows[r * w + c]; }
    }
}
static inline void to_rows(const uint32_t* cols, size_t h, size_t w, uint32_t* rows) {
    const size_t T = 512;
    long nt = (long)((h + T - 1) / T);
#pragma omp parallel for schedule(static) if (h * w > (1u << 16))
    for (long t = 0; t < nt; t++) {
        size_t r0 = (size_t)t * T, r1 = std::min(h, r0 + T);
        for (size_t c = 0; c < w; c++) { const uint32_t* in = cols + c * h; for (size_t r = r0; r < r1; r++) rows[r * w + c] = in[r]; }
    }
}

// Scratch that is NOT zero-filled by the allocating thread: for 100 MB buffers the page faults of a serial fill cost as
// much as the whole transform on 8 threads.
struct Scratch {
    std::unique_ptr<uint32_t[]> p;
    explicit Scratch(size_t n) : p(new uint32_t[n]) {}
    uint32_t* get() { return p.get(); }
    void reset() { p.reset(); }
};
// A zeroed h x w Matrix whose pages were first touched by all threads (the vector's own fill then runs at memset speed).
static inline Matrix matrix_touched_in_parallel(size_t h, size_t w) {
    Matrix m; m.width = w;
    const size_t n = h * w;
    m.v.reserve(n);
    volatile uint32_t* raw = m.v.data();
    const size_t STEP = 1024;                                       // one word per 4 KB page
#pragma omp parallel for schedule(static) if (n > (1u << 20))
    for (long i = 0; i < (long)((n + STEP - 1) / STEP); i++) raw[(size_t)i * STEP] = 0;
    m.v.resize(n);
    return m;
}

}  // namespace nttimpl

// In-place transform of every column, natural order in and out.
static inline void dft_rows(Matrix& m, bool inverse) {
    using namespace nttimpl;
    size_t h = m.height(), w = m.width;
    if (h <= 1) return;
    int lg = log2_strict(h);
    Plan p = make_plan(lg, inverse);
    std::vector<uint32_t> cols(h * w);
    to_columns(m.v.data(), h, w, cols.data());
    transform_columns(cols.data(), w, p, /*dif=*/true);             // natural in, bit-reversed out
    // back to row-major, undoing the bit reversal (and the 1/n of the inverse) on the way
    const uint32_t scale = inverse ? to_mont(inv((uint32_t)(h % P))) : to_mont(1);
#pragma omp parallel for schedule(static) if (h * w > (1u << 16))
    for (long i = 0; i < (long)h; i++) {
        size_t j = reverse_bits_len((uint32_t)i, lg);
        for (size_t c = 0; c < w; c++) m.v[(size_t)i * w + c] = mmul(cols[c * h + j], scale);
    }
}
static inline Matrix dft_batch(Matrix m) { dft_rows(m, false); return m; }
static inline Matrix idft_batch(Matrix m) { dft_rows(m, true); return m; }

// coset_lde_batch(mat, added_bits, shift): iDFT -> zero-pad -> coefficient i *= shift^i -> DFT.
// Evaluations of each column polynomial over shift * K, |K| = h << added_bits, natural order.
// The inverse runs as DIF (coefficients come out bit-reversed), the forward as DIT (takes them bit-reversed), so no
// permutation pass is needed in between; coefficient i of a column sits at index brev_h(i) << added_bits of the padded column.
static inline Matrix coset_lde_batch(const Matrix& in, int added_bits, uint32_t shift) {
    using namespace nttimpl;
    size_t h = in.height(), w = in.width, H = h << added_bits;
    if (h == 0 || w == 0) return Matrix(H, w);
    int lg = log2_strict(h);
    Scratch small_buf(h * w), big_buf(H * w);                       // first touched by the threads that fill them
    uint32_t* small = small_buf.get();
    uint32_t* big = big_buf.get();
    to_columns(in.v.data(), h, w, small);
    transform_columns(small, w, make_plan(lg, true), /*dif=*/true);
    std::vector<uint32_t> sp = geometric(1, shift, h), sc(h);
    const uint32_t ninv = inv((uint32_t)(h % P));
#pragma omp parallel for schedule(static) if (h > (1u << 14))
    for (long j = 0; j < (long)h; j++) sc[j] = to_mont(mul(sp[reverse_bits_len((uint32_t)j, lg)], ninv));
    const size_t CH = std::min<size_t>(h, 4096), nch = h / CH;
#pragma omp parallel for schedule(static) if (h * w > (1u << 16))
    for (long t = 0; t < (long)(w * nch); t++) {
        size_t c = (size_t)t / nch, j0 = ((size_t)t % nch) * CH;
        const uint32_t* a = small + c * h;
        uint32_t* o = big + c * H;
        const size_t pad = (size_t(1) << added_bits) - 1;
        for (size_t j = j0; j < j0 + CH; j++) {
            uint32_t* q = o + (j << added_bits);
            q[0] = mmul(a[j], sc[j]);
            for (size_t z = 1; z <= pad; z++) q[z] = 0;
        }
    }
    small_buf.reset();
    transform_columns(big, w, make_plan(lg + added_bits, false), /*dif=*/false);
    Matrix out = matrix_touched_in_parallel(H, w);
    to_rows(big, H, w, out.v.data());
    return out;
}

// Textbook O(n^2) evaluation used by the tests to pin dft_rows itself.
static inline Matrix naive_dft(const Matrix& in) {
    size_t h = in.height(), w = in.width;
    Matrix out(h, w);
    if (h == 0) return out;
    uint32_t root = two_adic_generator(log2_strict(h));
    for (size_t k = 0; k < h; k++) {
        uint32_t wk = pw(root, k);
        for (size_t c = 0; c < w; c++) {
            uint32_t acc = 0, x = 1;
            for (size_t j = 0; j < h; j++) { acc = add(acc, mul(in.v[j * w + c], x)); x = mul(x, wk); }
            out.v[k * w + c] = acc;
        }
    }
    return out;
}

}  // namespace orc
