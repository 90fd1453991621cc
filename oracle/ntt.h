// ORACLE — TEST INFRASTRUCTURE ONLY (see field.h header).
// TwoAdicSubgroupDft::{dft_batch, idft_batch, coset_lde_batch} of p3-dft, as reached from
// TwoAdicFriPcs::commit_shifted_batches (reference call sites derive/src/lib.rs:309,330,355,372;
// DFT type chosen at basic/src/bin/valida.rs:379 / basic/tests/test_prover.rs:436).
// Field arithmetic is exact, so Radix2DitParallel and Radix2Bowers give identical outputs;
// this restates the mathematical transform: out[k] = sum_j in[j] * w^(j k), w = two_adic_generator(log n),
// natural order in and out, per column of a row-major matrix.  [P3-UNVERIFIED; App. A item 4]
#pragma once
#include "field.h"

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

// In-place forward transform of every column (decimation in time, rows permuted first).
static inline void dft_rows(Matrix& m, bool inverse) {
    size_t h = m.height(), w = m.width;
    if (h <= 1) return;
    int lg = log2_strict(h);
    bit_reverse_rows(m);
    uint32_t root = two_adic_generator(lg);
    if (inverse) root = inv(root);
    for (int s = 1; s <= lg; s++) {
        size_t half = 1ull << (s - 1), len = half * 2;
        uint32_t wm = exp_pow2(root, lg - s);
        std::vector<uint32_t> tw = geometric(1, wm, half);
        long nb = (long)(h / len);
#pragma omp parallel for schedule(static) if (h * w > (1u << 16))
        for (long b = 0; b < nb * (long)half; b++) {
            size_t blk = (size_t)b / half, k = (size_t)b % half;
            uint32_t* x = m.v.data() + (blk * len + k) * w;
            uint32_t* y = x + half * w;
            uint32_t t = tw[k];
            for (size_t c = 0; c < w; c++) {
                uint32_t u = x[c], v = mul(y[c], t);
                x[c] = add(u, v);
                y[c] = sub(u, v);
            }
        }
    }
    if (inverse) {
        uint32_t ninv = inv((uint32_t)(h % P));
#pragma omp parallel for schedule(static) if (h * w > (1u << 16))
        for (long i = 0; i < (long)(h * w); i++) m.v[i] = mul(m.v[i], ninv);
    }
}
static inline Matrix dft_batch(Matrix m) { dft_rows(m, false); return m; }
static inline Matrix idft_batch(Matrix m) { dft_rows(m, true); return m; }

// coset_lde_batch(mat, added_bits, shift): iDFT -> zero-pad -> coefficient i *= shift^i -> DFT.
// Evaluations of each column polynomial over shift * K, |K| = h << added_bits, natural order.
static inline Matrix coset_lde_batch(const Matrix& in, int added_bits, uint32_t shift) {
    Matrix coeffs = idft_batch(in);
    size_t h = in.height(), w = in.width, H = h << added_bits;
    Matrix out(H, w);
    std::vector<uint32_t> sp = geometric(1, shift, h);
#pragma omp parallel for schedule(static) if (h * w > (1u << 16))
    for (long i = 0; i < (long)h; i++)
        for (size_t c = 0; c < w; c++) out.v[i * w + c] = mul(coeffs.v[i * w + c], sp[i]);
    dft_rows(out, false);
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
