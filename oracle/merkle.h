// ORACLE — TEST INFRASTRUCTURE ONLY (see field.h header).
// FieldMerkleTreeMmcs<BabyBear, SerializingHasher32<Keccak256Hash>,
//                     CompressionFunctionFromHasher<_, _, 2, 8>, 8>
// as configured at basic/src/bin/valida.rs:367-374 (tests: basic/tests/test_prover.rs:424-431).
// Restated from p3-merkle-tree / p3-symmetric [P3-UNVERIFIED; SURVEY App. A items 6-8]:
//  * leaf  = Keccak-256 over the concatenated rows (canonical u32, little-endian) of every matrix
//            of the maximal height (stable order), 32-byte output -> 8 words, each reduced mod p
//            (from_wrapped_u32);
//  * node  = hash of the 16 words of (left,right);
//  * when the layer length matches the (power-of-two) height of shorter matrices:
//            node = compress(compress(left,right), hash(rows of those matrices)).
#pragma once
#include "field.h"
#include "keccak.h"
#include <algorithm>
#include <numeric>

namespace orc {

using Digest = std::array<uint32_t, 8>;

static inline Digest hash_words(const uint32_t* const* slices, const size_t* lens, size_t nslices) {
    size_t total = 0;
    for (size_t i = 0; i < nslices; i++) total += lens[i];
    std::vector<uint8_t> bytes(total * 4);
    size_t o = 0;
    for (size_t i = 0; i < nslices; i++)
        for (size_t k = 0; k < lens[i]; k++) {
            uint32_t x = slices[i][k];
            bytes[o++] = (uint8_t)x; bytes[o++] = (uint8_t)(x >> 8); bytes[o++] = (uint8_t)(x >> 16); bytes[o++] = (uint8_t)(x >> 24);
        }
    uint8_t out[32];
    keccak256(bytes.data(), bytes.size(), out);
    Digest d;
    for (int i = 0; i < 8; i++) {
        uint32_t w = (uint32_t)out[4 * i] | ((uint32_t)out[4 * i + 1] << 8) | ((uint32_t)out[4 * i + 2] << 16) | ((uint32_t)out[4 * i + 3] << 24);
        d[i] = w % P;
    }
    return d;
}
static inline Digest hash_slice(const uint32_t* p, size_t n) { return hash_words(&p, &n, 1); }
static inline Digest compress2(const Digest& l, const Digest& r) {
    uint32_t buf[16];
    std::memcpy(buf, l.data(), 32);
    std::memcpy(buf + 8, r.data(), 32);
    return hash_slice(buf, 16);
}

struct MerkleTree {
    std::vector<Matrix> leaves;                  // in the caller's order
    std::vector<std::vector<Digest>> layers;     // layers[0] = leaf digests ... back() = {root}
    Digest root() const { return layers.back()[0]; }
    size_t max_height() const { size_t m = 0; for (auto& l : leaves) m = std::max(m, l.height()); return m; }
};

static inline Digest hash_rows(const std::vector<const Matrix*>& mats, size_t row) {
    std::vector<const uint32_t*> sl(mats.size());
    std::vector<size_t> ln(mats.size());
    for (size_t i = 0; i < mats.size(); i++) { sl[i] = mats[i]->row(row); ln[i] = mats[i]->width; }
    return hash_words(sl.data(), ln.data(), mats.size());
}

static inline MerkleTree merkle_commit(std::vector<Matrix> leaves) {
    MerkleTree t;
    t.leaves = std::move(leaves);
    assert(!t.leaves.empty());
    std::vector<size_t> order(t.leaves.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return t.leaves[a].height() > t.leaves[b].height(); });
    size_t pos = 0;
    size_t max_h = t.leaves[order[0]].height();
    std::vector<const Matrix*> tallest;
    while (pos < order.size() && t.leaves[order[pos]].height() == max_h) tallest.push_back(&t.leaves[order[pos++]]);
    std::vector<Digest> first(max_h);
#pragma omp parallel for schedule(static) if (max_h > 256)
    for (long i = 0; i < (long)max_h; i++) first[i] = hash_rows(tallest, (size_t)i);
    t.layers.push_back(std::move(first));
    while (t.layers.back().size() > 1) {
        const std::vector<Digest>& prev = t.layers.back();
        size_t next_len = prev.size() / 2;
        std::vector<const Matrix*> inject;
        while (pos < order.size() && (1ull << log2_ceil(t.leaves[order[pos]].height())) == next_len) inject.push_back(&t.leaves[order[pos++]]);
        std::vector<Digest> next(next_len);
#pragma omp parallel for schedule(static) if (next_len > 256)
        for (long i = 0; i < (long)next_len; i++) {
            Digest d = compress2(prev[2 * i], prev[2 * i + 1]);
            if (!inject.empty()) d = compress2(d, hash_rows(inject, (size_t)i));
            next[i] = d;
        }
        t.layers.push_back(std::move(next));
    }
    assert(pos == order.size());
    return t;
}

struct BatchOpening {
    std::vector<std::vector<uint32_t>> opened_values;  // one row per matrix, caller's order
    std::vector<Digest> opening_proof;                 // siblings, leaf level first
};

static inline BatchOpening merkle_open(const MerkleTree& t, size_t index) {
    BatchOpening o;
    int log_max = log2_ceil(t.max_height());
    for (auto& m : t.leaves) {
        int bits_reduced = log_max - log2_ceil(m.height());
        size_t r = index >> bits_reduced;
        o.opened_values.emplace_back(m.row(r), m.row(r) + m.width);
    }
    for (int i = 0; i < log_max; i++) o.opening_proof.push_back(t.layers[i][(index >> i) ^ 1]);
    return o;
}

struct Dims { size_t width, height; };

static inline bool merkle_verify(const Digest& commit, const std::vector<Dims>& dims, size_t index,
                                 const std::vector<std::vector<uint32_t>>& opened, const std::vector<Digest>& proof) {
    if (dims.size() != opened.size() || dims.empty()) return false;
    for (size_t i = 0; i < dims.size(); i++) if (opened[i].size() != dims[i].width) return false;
    std::vector<size_t> order(dims.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return dims[a].height > dims[b].height; });
    size_t pos = 0;
    size_t cur = 1ull << log2_ceil(dims[order[0]].height);
    if (proof.size() != (size_t)log2_ceil(cur)) return false;
    auto hash_group = [&](size_t padded) {
        std::vector<const uint32_t*> sl; std::vector<size_t> ln;
        while (pos < order.size() && (1ull << log2_ceil(dims[order[pos]].height)) == padded) {
            sl.push_back(opened[order[pos]].data()); ln.push_back(opened[order[pos]].size()); pos++;
        }
        return hash_words(sl.data(), ln.data(), sl.size());
    };
    Digest root = hash_group(cur);
    for (const Digest& sib : proof) {
        root = (index & 1) ? compress2(sib, root) : compress2(root, sib);
        index >>= 1;
        cur >>= 1;
        if (pos < order.size() && (1ull << log2_ceil(dims[order[pos]].height)) == cur) root = compress2(root, hash_group(cur));
    }
    return pos == order.size() && root == commit;
}

}  // namespace orc
