// ORACLE — TEST INFRASTRUCTURE ONLY (see field.h header).  C entry points for ctypes.
#include "field.h"
#include "keccak.h"
#include "poseidon.h"
#include "ntt.h"
#include "merkle.h"
#include "pcs.h"
#include "machine.h"

using namespace orc;

#include <omp.h>
#include <malloc.h>

extern "C" {

void orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
// For the process that TIMES the oracle as the CPU baseline (bench.py --impl reference): keep freed blocks in the heap instead
// of returning every multi-megabyte vector to the kernel and faulting its pages in again on the next allocation (what a
// Rust prover gets from its allocator); process-wide, so the test suite does not call it.  Returns 1 when all three took.
int orc_tune_allocator() {
    int ok = mallopt(M_MMAP_MAX, 0);
    ok &= mallopt(M_TRIM_THRESHOLD, -1) & mallopt(M_TOP_PAD, 256 << 20);    // -1: the heap is never trimmed
    return ok;
}
// After orc_tune_allocator: grow the heap to `bytes` and touch every page on all threads, then free it all — the pages stay
// with the process, so a proof that follows finds its working set already mapped.  A prover timed ONCE at a size (the
// full-workload proof of the reference arm) otherwise spends its time in first-touch page faults (7 M of them for the
// 2^22-row workload; tens of microseconds each on a virtualised host).  Returns the bytes it managed to touch.
uint64_t orc_prefault_heap(uint64_t bytes) {
    const uint64_t chunk = 1ull << 30;
    std::vector<char*> held;
    uint64_t done = 0;
    while (done < bytes) {
        const uint64_t n = bytes - done < chunk ? bytes - done : chunk;
        char* p = (char*)malloc(n);
        if (!p) break;
#pragma omp parallel for schedule(static)
        for (long long off = 0; off < (long long)n; off += 4096) p[off] = 1;
        held.push_back(p);
        done += n;
    }
    for (auto it = held.rbegin(); it != held.rend(); ++it) free(*it);
    return done;
}
int orc_max_threads() { return omp_get_max_threads(); }

void orc_keccak256(const uint8_t* in, uint64_t len, uint8_t* out, uint32_t pad) { keccak256_pad(in, len, out, (uint8_t)pad); }

uint32_t orc_two_adic_generator(uint32_t bits) { return two_adic_generator((int)bits); }
uint32_t orc_mul(uint32_t a, uint32_t b) { return mul(a, b); }
uint32_t orc_inv(uint32_t a) { return inv(a); }
void orc_ext_mul(const uint32_t* a, const uint32_t* b, uint32_t* o) { Ext5 x, y; memcpy(x.c, a, 20); memcpy(y.c, b, 20); Ext5 r = x * y; memcpy(o, r.c, 20); }
void orc_ext_inv(const uint32_t* a, uint32_t* o) { Ext5 x; memcpy(x.c, a, 20); Ext5 r = ext_inv(x); memcpy(o, r.c, 20); }

// row-major h x w, in place
void orc_dft(uint32_t* vals, uint64_t h, uint64_t w, int inverse) {
    Matrix m(std::vector<uint32_t>(vals, vals + h * w), w);
    dft_rows(m, inverse != 0);
    memcpy(vals, m.v.data(), h * w * 4);
}
void orc_naive_dft(const uint32_t* vals, uint64_t h, uint64_t w, uint32_t* out) {
    Matrix m(std::vector<uint32_t>(vals, vals + h * w), w);
    Matrix o = naive_dft(m);
    memcpy(out, o.v.data(), h * w * 4);
}
// out: (h << added_bits) x w row-major; bitrev != 0 -> rows in bit-reversed order (as committed)
void orc_coset_lde(const uint32_t* vals, uint64_t h, uint64_t w, uint32_t added_bits, uint32_t shift, int bitrev, uint32_t* out) {
    Matrix m(std::vector<uint32_t>(vals, vals + h * w), w);
    Matrix o = coset_lde_batch(m, (int)added_bits, shift);
    if (bitrev) bit_reverse_rows(o);
    memcpy(out, o.v.data(), o.v.size() * 4);
}

void orc_merkle_root(uint32_t n, const uint32_t* const* mats, const uint64_t* heights, const uint64_t* widths, uint32_t* digest) {
    std::vector<Matrix> ms;
    for (uint32_t i = 0; i < n; i++) ms.emplace_back(std::vector<uint32_t>(mats[i], mats[i] + heights[i] * widths[i]), widths[i]);
    MerkleTree t = merkle_commit(std::move(ms));
    Digest r = t.root();
    memcpy(digest, r.data(), 32);
}

// TwoAdicFriPcs::commit_batches / commit_shifted_batches: returns the root; if lde_out != NULL,
// lde_out[i] receives the committed (bit-reversed) LDE of matrix i, row-major (2h x w).
void orc_commit_batches(uint32_t n, const uint32_t* const* mats, const uint64_t* heights, const uint64_t* widths,
                        const uint32_t* coset_shifts_or_null, uint32_t* digest, uint32_t* const* lde_out) {
    std::vector<Matrix> ms;
    std::vector<uint32_t> shifts;
    for (uint32_t i = 0; i < n; i++) {
        ms.emplace_back(std::vector<uint32_t>(mats[i], mats[i] + heights[i] * widths[i]), widths[i]);
        shifts.push_back(coset_shifts_or_null ? coset_shifts_or_null[i] : 1);
    }
    Pcs pcs;
    PcsData d = pcs.commit_shifted_batches(ms, shifts);
    Digest r = d.tree.root();
    memcpy(digest, r.data(), 32);
    if (lde_out)
        for (uint32_t i = 0; i < n; i++)
            if (lde_out[i]) memcpy(lde_out[i], d.tree.leaves[i].v.data(), d.tree.leaves[i].v.size() * 4);
}

void orc_default_round_constants(uint32_t* out) { default_round_constants(out); }
void orc_coset_mds_matrix(uint32_t* out256) {
    Poseidon16 p; p.set_default_mds();
    for (int i = 0; i < 16; i++) for (int j = 0; j < 16; j++) out256[i * 16 + j] = p.mds[i][j];
}
void orc_poseidon_permute(const uint32_t* rc480, uint32_t* state16) {
    Poseidon16 p; memcpy(p.rc, rc480, sizeof p.rc); p.set_default_mds();
    p.permute(state16);
}
// Challenger script: ops[i] = 0 observe(args[i]) | 1 sample() -> out | 2 sample_bits(args[i]) -> out | 3 grind(args[i]) -> out
void orc_challenger_script(const uint32_t* rc480, uint32_t n, const uint32_t* ops, const uint32_t* args, uint32_t* out) {
    Poseidon16 p; memcpy(p.rc, rc480, sizeof p.rc); p.set_default_mds();
    Challenger ch(&p);
    for (uint32_t i = 0; i < n; i++) {
        switch (ops[i]) {
            case 0: ch.observe(args[i]); out[i] = 0; break;
            case 1: out[i] = ch.sample(); break;
            case 2: out[i] = ch.sample_bits((int)args[i]); break;
            case 3: out[i] = ch.grind((int)args[i]); break;
        }
    }
}

}  // extern "C"

#include "machine_api.inc"
