// Machine::verify for proofs in the reference's wire format — the acceptance side of the boundary.
// Follows verify() (derive/src/lib.rs:492-650; hand copy basic/src/lib.rs:677-840):
//   re-commit the preprocessed traces (device, same kernels as the prover) -> replay the transcript ->
//   TwoAdicFriPcs::verify_multi_batches + p3-fri verify_query [P3-UNVERIFIED; SURVEY App. A items 14, 15]
//   -> per-chip verify_constraints (verify.cu) -> the cumulative sums of all chips add to zero.
// Everything except the preprocessed commit is host arithmetic on a ~2 MB proof (40 queries x ~25 Merkle
// paths); it exists so that a caller of this library can check what it produced without the Rust
// verifier, and so that the tests can cross-check prover and verifier against the oracle in both directions.
#include "../ctx.h"
#include "../verify.h"
#include "challenger.h"
#include <array>
#include <cstring>
#include <string>

using bb::E5;

namespace {

constexpr int LOG_BLOWUP = 1, NUM_QUERIES = 40, POW_BITS = 8;   // basic/src/bin/valida.rs:385-390
constexpr int MAX_LOG_DEGREE = 26;

using Digest = std::array<uint32_t, 8>;   // canonical words

// ---- Keccak-256 on the host (original 0x01 padding: p3-keccak wraps tiny-keccak's Keccak::v256) -------
const uint64_t RC[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull,
                         0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull,
                         0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull,
                         0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
                         0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
inline uint64_t rotl(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
void keccak_f(uint64_t a[25]) {
    for (int round = 0; round < 24; round++) {
        uint64_t c[5], d[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
        // rho + pi walk: lane (x, y) moves to (y, 2x + 3y) with rotation (t+1)(t+2)/2
        int x = 1, y = 0;
        b[0] = a[0];
        for (int t = 0; t < 24; t++) {
            const int nx = y, ny = (2 * x + 3 * y) % 5;
            b[nx + 5 * ny] = rotl(a[x + 5 * y], ((t + 1) * (t + 2) / 2) % 64);
            x = nx; y = ny;
        }
        for (int yy = 0; yy < 25; yy += 5)
            for (int xx = 0; xx < 5; xx++) a[yy + xx] = b[yy + xx] ^ (~b[yy + (xx + 1) % 5] & b[yy + (xx + 2) % 5]);
        a[0] ^= RC[round];
    }
}
// SerializingHasher32<Keccak256Hash>: little-endian canonical words in, 32 bytes out -> 8 words, each reduced mod p
Digest hash_canonical_words(const std::vector<uint32_t>& w) {
    uint64_t st[25] = {0};
    const size_t rate_words = 34;   // 136 bytes
    size_t i = 0;
    while (w.size() - i >= rate_words) {
        for (size_t k = 0; k < 17; k++) st[k] ^= (uint64_t)w[i + 2 * k] | ((uint64_t)w[i + 2 * k + 1] << 32);
        keccak_f(st);
        i += rate_words;
    }
    uint8_t block[136] = {0};
    size_t rem = w.size() - i;
    for (size_t k = 0; k < rem; k++) for (int b = 0; b < 4; b++) block[4 * k + b] = (uint8_t)(w[i + k] >> (8 * b));
    block[4 * rem] ^= 0x01;
    block[135] ^= 0x80;
    for (size_t k = 0; k < 17; k++) { uint64_t v = 0; for (int b = 7; b >= 0; b--) v = (v << 8) | block[8 * k + b]; st[k] ^= v; }
    keccak_f(st);
    Digest d;
    for (int k = 0; k < 4; k++) { d[2 * k] = (uint32_t)st[k] % bb::P; d[2 * k + 1] = (uint32_t)(st[k] >> 32) % bb::P; }
    return d;
}
Digest compress2(const Digest& l, const Digest& r) {
    std::vector<uint32_t> w(16);
    std::memcpy(w.data(), l.data(), 32); std::memcpy(w.data() + 8, r.data(), 32);
    return hash_canonical_words(w);
}

int log2_ceil(uint64_t n) { int l = 0; while ((1ull << l) < n) l++; return l; }

// FieldMerkleTreeMmcs::verify_batch [P3-UNVERIFIED; SURVEY App. A item 8]: matrices sorted by height (stable, tallest
// first); rows of equal padded height are hashed together; a shorter group is injected when the running height reaches it.
struct Dim { uint64_t w, h; };
bool merkle_verify_batch(const Digest& commit, const std::vector<Dim>& dims, uint64_t index,
                         const std::vector<std::vector<uint32_t>>& opened_canonical, const std::vector<Digest>& path) {
    if (dims.empty() || dims.size() != opened_canonical.size()) return false;
    std::vector<size_t> order;
    for (size_t i = 0; i < dims.size(); i++) { if (opened_canonical[i].size() != dims[i].w) return false; order.push_back(i); }
    for (size_t i = 1; i < order.size(); i++)   // stable insertion sort, descending height
        for (size_t j = i; j > 0 && dims[order[j - 1]].h < dims[order[j]].h; j--) std::swap(order[j - 1], order[j]);
    size_t pos = 0;
    int level = log2_ceil(dims[order[0]].h);
    if (path.size() != (size_t)level) return false;
    auto group = [&](int lvl) {
        std::vector<uint32_t> cat;
        while (pos < order.size() && log2_ceil(dims[order[pos]].h) == lvl) { auto& r = opened_canonical[order[pos]]; cat.insert(cat.end(), r.begin(), r.end()); pos++; }
        return hash_canonical_words(cat);
    };
    Digest node = group(level);
    for (const Digest& sib : path) {
        node = (index & 1) ? compress2(sib, node) : compress2(node, sib);
        index >>= 1; level--;
        if (pos < order.size() && log2_ceil(dims[order[pos]].h) == level) node = compress2(node, group(level));
    }
    return pos == order.size() && node == commit;
}

// ---- CBOR reader for exactly the shape vgpu_prove / the reference's ciborium writer emit ----------------
struct Reader {
    const uint8_t* p; const uint8_t* end; bool ok = true;
    uint64_t head(int major) {
        if (!ok || p >= end) { ok = false; return 0; }
        uint8_t b = *p++;
        if ((b >> 5) != major) { ok = false; return 0; }
        uint8_t info = b & 31;
        if (info < 24) return info;
        int n = info == 24 ? 1 : info == 25 ? 2 : info == 26 ? 4 : info == 27 ? 8 : -1;
        if (n < 0 || end - p < n) { ok = false; return 0; }
        uint64_t v = 0;
        for (int i = 0; i < n; i++) v = (v << 8) | *p++;
        return v;
    }
    void key(const char* s) {
        uint64_t n = head(3), want = std::strlen(s);
        if (!ok || n != want || (uint64_t)(end - p) < n || std::memcmp(p, s, n) != 0) { ok = false; return; }
        p += n;
    }
    void map(uint64_t n) { if (head(5) != n) ok = false; }
    // bounded array length: every element costs at least one byte, so a hostile length cannot make us allocate
    uint64_t arr() { uint64_t n = head(4); if (n > (uint64_t)(end - p)) { ok = false; return 0; } return n; }
    uint32_t felt() {   // BabyBear { value: Montgomery word }
        map(1); key("value");
        uint64_t v = head(0);
        if (v >= bb::P) ok = false;
        return (uint32_t)v;
    }
    E5 ext() { E5 e = bb::e5_zero(); map(1); key("value"); if (arr() != 5) ok = false; for (int i = 0; i < 5 && ok; i++) e.c[i] = felt(); return e; }
    Digest digest() { Digest d{}; if (arr() != 8) ok = false; for (int i = 0; i < 8 && ok; i++) d[i] = bb::from_monty(felt()); return d; }
    std::vector<Digest> digests() { std::vector<Digest> v; uint64_t n = arr(); for (uint64_t i = 0; i < n && ok; i++) v.push_back(digest()); return v; }
    std::vector<E5> exts() { std::vector<E5> v; uint64_t n = arr(); for (uint64_t i = 0; i < n && ok; i++) v.push_back(ext()); return v; }
};

struct BatchOpeningV { std::vector<std::vector<uint32_t>> rows_monty; std::vector<Digest> path; };
struct FriStepV { E5 sibling; std::vector<Digest> path; };
struct ProofV {
    Digest main_commit, perm_commit, quot_commit;
    std::vector<Digest> fri_commits;
    std::vector<std::vector<FriStepV>> fri_queries;
    E5 final_poly; uint32_t pow_witness_monty = 0;
    std::vector<std::vector<BatchOpeningV>> query_openings;   // [query][round]
    struct Chip { uint32_t log_degree = 0; VgChipOpening ov; size_t n_prep_local = 0, n_prep_next = 0; E5 cumulative_sum; };
    std::vector<Chip> chips;
};

bool decode(const uint8_t* data, uint64_t len, ProofV* out) {
    Reader r{data, data + len};
    r.map(3);
    r.key("commitments"); r.map(3);
    r.key("main_trace"); out->main_commit = r.digest();
    r.key("perm_trace"); out->perm_commit = r.digest();
    r.key("quotient_chunks"); out->quot_commit = r.digest();
    r.key("opening_proof"); r.map(2);
    r.key("fri_proof"); r.map(4);
    r.key("commit_phase_commits"); out->fri_commits = r.digests();
    r.key("query_proofs");
    for (uint64_t q = 0, nq = r.arr(); q < nq && r.ok; q++) {
        r.map(1); r.key("commit_phase_openings");
        std::vector<FriStepV> steps;
        for (uint64_t s = 0, ns = r.arr(); s < ns && r.ok; s++) {
            FriStepV st;
            r.map(2); r.key("sibling_value"); st.sibling = r.ext(); r.key("opening_proof"); st.path = r.digests();
            steps.push_back(std::move(st));
        }
        out->fri_queries.push_back(std::move(steps));
    }
    r.key("final_poly"); out->final_poly = r.ext();
    r.key("pow_witness"); out->pow_witness_monty = r.felt();
    r.key("query_openings");
    for (uint64_t q = 0, nq = r.arr(); q < nq && r.ok; q++) {
        std::vector<BatchOpeningV> per_round;
        for (uint64_t b = 0, nb = r.arr(); b < nb && r.ok; b++) {
            BatchOpeningV bo;
            r.map(2); r.key("opened_values");
            for (uint64_t m = 0, nm = r.arr(); m < nm && r.ok; m++) {
                std::vector<uint32_t> row;
                for (uint64_t c = 0, nc = r.arr(); c < nc && r.ok; c++) row.push_back(r.felt());
                bo.rows_monty.push_back(std::move(row));
            }
            r.key("opening_proof"); bo.path = r.digests();
            per_round.push_back(std::move(bo));
        }
        out->query_openings.push_back(std::move(per_round));
    }
    r.key("chip_proofs");
    for (uint64_t i = 0, n = r.arr(); i < n && r.ok; i++) {
        ProofV::Chip c;
        r.map(3);
        r.key("log_degree"); c.log_degree = (uint32_t)r.head(0);
        r.key("opened_values"); r.map(7);
        r.key("preprocessed_local"); c.n_prep_local = r.exts().size();
        r.key("preprocessed_next"); c.n_prep_next = r.exts().size();
        r.key("trace_local"); c.ov.trace_local = r.exts();
        r.key("trace_next"); c.ov.trace_next = r.exts();
        r.key("permutation_local"); c.ov.perm_local = r.exts();
        r.key("permutation_next"); c.ov.perm_next = r.exts();
        r.key("quotient_chunks"); c.ov.quotient_chunks = r.exts();
        r.key("cumulative_sum"); c.cumulative_sum = r.ext();
        out->chips.push_back(std::move(c));
    }
    return r.ok && r.p == r.end;
}

struct RoundV { Digest commit; std::vector<Dim> dims; std::vector<std::vector<E5>> points; std::vector<std::vector<const std::vector<E5>*>> values; };

// TwoAdicFriPcs::verify_multi_batches + p3-fri verifier; 0 = accept, otherwise the verdict code of include/valida_b200.h
int32_t verify_openings(const std::vector<RoundV>& rounds, const ProofV& pf, vgh::Challenger& ch) {
    const E5 alpha = ch.sample_ext();
    std::vector<E5> betas;
    for (const Digest& c : pf.fri_commits) { ch.observe_digest_canonical(c.data()); betas.push_back(ch.sample_ext()); }
    if (pf.fri_queries.size() != (size_t)NUM_QUERIES || pf.query_openings.size() != (size_t)NUM_QUERIES) return VGPU_REJECT_SHAPE;
    if (!ch.check_witness(POW_BITS, pf.pow_witness_monty)) return VGPU_REJECT_POW;
    const int log_max_height = (int)pf.fri_commits.size() + LOG_BLOWUP;
    if (log_max_height > MAX_LOG_DEGREE + LOG_BLOWUP) return VGPU_REJECT_SHAPE;
    for (auto& rd : rounds)
        for (auto& d : rd.dims) if (log2_ceil(d.h) + LOG_BLOWUP > log_max_height) return VGPU_REJECT_SHAPE;
    std::vector<uint32_t> indices;
    for (int q = 0; q < NUM_QUERIES; q++) indices.push_back(ch.sample_bits(log_max_height));
    const uint32_t gen = bb::to_monty(bb::GEN_CANON);
    for (int q = 0; q < NUM_QUERIES; q++) {
        uint64_t index = indices[q];
        E5 ro[32], apw[32];
        for (int i = 0; i < 32; i++) { ro[i] = bb::e5_zero(); apw[i] = bb::e5_one(); }
        if (pf.query_openings[q].size() != rounds.size()) return VGPU_REJECT_SHAPE;
        for (size_t r = 0; r < rounds.size(); r++) {
            const RoundV& rd = rounds[r];
            const BatchOpeningV& bo = pf.query_openings[q][r];
            if (bo.rows_monty.size() != rd.dims.size()) return VGPU_REJECT_SHAPE;
            std::vector<Dim> lde_dims;
            uint64_t max_h = 0;
            for (auto& d : rd.dims) { lde_dims.push_back({d.w, d.h << LOG_BLOWUP}); max_h = std::max(max_h, d.h << LOG_BLOWUP); }
            std::vector<std::vector<uint32_t>> canon_rows;
            for (auto& row : bo.rows_monty) { std::vector<uint32_t> c; for (uint32_t x : row) c.push_back(bb::from_monty(x)); canon_rows.push_back(std::move(c)); }
            const uint64_t batch_index = index >> (log_max_height - log2_ceil(max_h));
            if (!merkle_verify_batch(rd.commit, lde_dims, batch_index, canon_rows, bo.path)) return VGPU_REJECT_INPUT_MERKLE;
            for (size_t mi = 0; mi < rd.dims.size(); mi++) {
                const int lh = log2_ceil(rd.dims[mi].h) + LOG_BLOWUP;
                const uint32_t rev = bb::reverse_bits((uint32_t)(index >> (log_max_height - lh)), lh);
                const uint32_t x = bb::mul(gen, bb::pow(bb::two_adic_generator_monty(lh), rev));
                for (size_t pi = 0; pi < rd.points[mi].size(); pi++) {
                    const std::vector<E5>& at_z = *rd.values[mi][pi];
                    if (at_z.size() != bo.rows_monty[mi].size()) return VGPU_REJECT_SHAPE;
                    const E5 den = bb::e5_add_base(bb::e5_neg(rd.points[mi][pi]), x);   // x - z
                    if (bb::e5_is_zero(den)) return VGPU_REJECT_SHAPE;
                    const E5 dinv = bb::e5_inv(den);
                    for (size_t c = 0; c < at_z.size(); c++) {
                        const E5 quotient = bb::e5_mul(bb::e5_add_base(bb::e5_neg(at_z[c]), bo.rows_monty[mi][c]), dinv);   // (p(x) - p(z)) / (x - z)
                        ro[lh] = bb::e5_add(ro[lh], bb::e5_mul(apw[lh], quotient));
                        apw[lh] = bb::e5_mul(apw[lh], alpha);
                    }
                }
            }
        }
        // p3-fri verify_query
        const std::vector<FriStepV>& steps = pf.fri_queries[q];
        if (steps.size() != pf.fri_commits.size()) return VGPU_REJECT_SHAPE;
        E5 folded = bb::e5_zero();
        uint32_t x = bb::pow(bb::two_adic_generator_monty(log_max_height), bb::reverse_bits((uint32_t)index, log_max_height));
        const uint32_t minus_one = bb::two_adic_generator_monty(1);
        size_t si = 0;
        for (int lfh = log_max_height - 1; lfh >= LOG_BLOWUP; lfh--, si++) {
            folded = bb::e5_add(folded, ro[lfh + 1]);
            const uint64_t sib = (index ^ 1) & 1, pair = index >> 1;
            E5 evals[2] = {folded, folded};
            evals[sib] = steps[si].sibling;
            std::vector<uint32_t> row(10);
            for (int e = 0; e < 2; e++) for (int l = 0; l < 5; l++) row[5 * e + l] = bb::from_monty(evals[e].c[l]);
            if (!merkle_verify_batch(pf.fri_commits[si], {{10, 1ull << lfh}}, pair, {row}, steps[si].path)) return VGPU_REJECT_FRI_MERKLE;
            uint32_t xs[2] = {x, x};
            xs[sib] = bb::mul(xs[sib], minus_one);
            // line through (xs[0], evals[0]), (xs[1], evals[1]) evaluated at beta; xs[1] - xs[0] = -2 xs[0]
            const uint32_t slope_den = bb::inv(bb::sub(xs[1], xs[0]));
            const E5 slope = bb::e5_mul_base(bb::e5_sub(evals[1], evals[0]), slope_den);
            folded = bb::e5_add(evals[0], bb::e5_mul(bb::e5_sub_base(betas[si], xs[0]), slope));
            index = pair;
            x = bb::sqr(x);
        }
        // The prover's last fold also adds the reduced openings of the height-2 LDEs (traces of ONE row: constant polynomials,
        // for which (p(x) - p(z)) / (x - z) is exactly 0).  The loop above never reaches them, so they are checked here: without
        // this the opened values of every one-row chip — and with them its cumulative sum — would be bound by nothing.
        if (!bb::e5_is_zero(ro[LOG_BLOWUP])) return VGPU_REJECT_FRI_FINAL;
        for (int l = 0; l < 5; l++) if (folded.c[l] != pf.final_poly.c[l]) return VGPU_REJECT_FRI_FINAL;
    }
    return VGPU_ACCEPT;
}

}  // namespace

extern "C" int32_t vgpu_verify(vgpu_ctx* ctx, const uint8_t* proof, uint64_t proof_len, const vgpu_matrix prep[2], int32_t repr, int32_t* verdict) {
    if (!ctx) return -1;
    if (!proof || !prep || !verdict) VG_FAIL(ctx, "verify: null argument");
    if (!ctx->challenger_set) VG_FAIL(ctx, "verify: vgpu_set_challenger has not been called");
    VG_TRY(vg_enter(ctx));
    *verdict = VGPU_REJECT_MALFORMED;
    ProofV pf;
    if (!decode(proof, proof_len, &pf)) return 0;
    if (pf.chips.size() != (size_t)VGPU_NUM_CHIPS) { *verdict = VGPU_REJECT_SHAPE; return 0; }
    for (auto& c : pf.chips) if (c.log_degree > (uint32_t)MAX_LOG_DEGREE || c.n_prep_local || c.n_prep_next) { *verdict = VGPU_REJECT_SHAPE; return 0; }
    // the two chips with preprocessed columns have the height of those columns (program ROM, range table): a proof may not
    // shrink them (an all-one-row proof has no FRI layers at all)
    if ((1ull << pf.chips[1].log_degree) != prep[0].height || (1ull << pf.chips[12].log_degree) != prep[1].height) { *verdict = VGPU_REJECT_SHAPE; return 0; }

    vgh::Poseidon16 perm;
    perm.set(ctx->poseidon_rc, ctx->poseidon_has_mds ? ctx->poseidon_mds : nullptr);
    vgh::Challenger ch;
    ch.perm = &perm;
    {   // preprocessed commitment, recomputed (derive/src/lib.rs:505-517)
        uint32_t digest[8];
        vgpu_prover_data* pd = nullptr;
        // a verifier checks alone: no collective here even when the context is a rank of a split prover
        const bool was_sharding = ctx->sharding;
        ctx->sharding = false;
        const int32_t rc = vgpu_commit_batches_host(ctx, prep, 2, repr, nullptr, digest, &pd);
        ctx->sharding = was_sharding;
        if (rc) return rc;
        vgpu_prover_data_free(pd);
        ch.observe_digest_canonical(digest);
    }
    ch.observe_digest_canonical(pf.main_commit.data());
    uint32_t perm_challenges[15];
    for (int i = 0; i < 3; i++) { E5 e = ch.sample_ext(); for (int l = 0; l < 5; l++) perm_challenges[5 * i + l] = bb::from_monty(e.c[l]); }
    ch.observe_digest_canonical(pf.perm_commit.data());
    const E5 alpha = ch.sample_ext();
    ch.observe_digest_canonical(pf.quot_commit.data());
    const E5 zeta = ch.sample_ext();

    std::vector<RoundV> rounds(3);
    rounds[0].commit = pf.main_commit; rounds[1].commit = pf.perm_commit; rounds[2].commit = pf.quot_commit;
    for (int i = 0; i < VGPU_NUM_CHIPS; i++) {
        const vgpu_chip_desc* chip = vgpu_basic_machine_chip(i);
        const ProofV::Chip& c = pf.chips[i];
        const uint64_t h = 1ull << c.log_degree;
        const E5 zg = bb::e5_mul_base(zeta, bb::two_adic_generator_monty((int)c.log_degree));
        rounds[0].dims.push_back({chip->width, h});
        rounds[1].dims.push_back({5ull * (chip->n_interactions + 1), h});
        rounds[2].dims.push_back({10, h});
        rounds[0].points.push_back({zeta, zg}); rounds[0].values.push_back({&c.ov.trace_local, &c.ov.trace_next});
        rounds[1].points.push_back({zeta, zg}); rounds[1].values.push_back({&c.ov.perm_local, &c.ov.perm_next});
        rounds[2].points.push_back({bb::e5_sqr(zeta)}); rounds[2].values.push_back({&c.ov.quotient_chunks});
    }
    int32_t v = verify_openings(rounds, pf, ch);
    if (v != VGPU_ACCEPT) { *verdict = v; return 0; }
    for (int i = 0; i < VGPU_NUM_CHIPS; i++) {
        bool ok = false;
        VG_TRY(vg_verify_chip_constraints(ctx, vgpu_basic_machine_chip(i), pf.chips[i].log_degree, pf.chips[i].ov, pf.chips[i].cumulative_sum,
                                          zeta, alpha, perm_challenges, &ok));
        if (!ok) { *verdict = VGPU_REJECT_CONSTRAINTS_CHIP0 - i; return 0; }
    }
    E5 sum = bb::e5_zero();
    for (auto& c : pf.chips) sum = bb::e5_add(sum, c.cumulative_sum);
    if (!bb::e5_is_zero(sum)) { *verdict = VGPU_REJECT_CUMULATIVE_SUM; return 0; }
    *verdict = VGPU_ACCEPT;
    return 0;
}
