// Machine::prove on the device — orchestration, transcript and proof assembly.
// Mirrors, step for step, the reference's prove() (derive/src/lib.rs:275-446; hand copy
// basic/src/lib.rs:147-675) and TwoAdicFriPcs::open_multi_batches / p3-fri prove
// [P3-UNVERIFIED; SURVEY App. A items 14, 15]; every observe/sample occurs in the reference's order.
// Heavy steps are the CUDA kernels of ntt.cu / merkle.cu / perm.cu / quotient.cu / open.cu; this file
// holds no field loops over trace-sized data.
#include "../ctx.h"
#include "../merkle.h"
#include "../open.h"
#include "../devchip.h"
#include "challenger.h"
#include <algorithm>
#include <array>
#include <chrono>
#include <cstring>
#include <map>
#include <memory>

using bb::E5;

namespace {

constexpr int LOG_BLOWUP = 1, NUM_QUERIES = 40, POW_BITS = 8;   // basic/src/bin/valida.rs:385-390

using Digest = std::array<uint32_t, 8>;   // canonical words
struct ExtC { uint32_t c[5]; };           // canonical
ExtC canon(const E5& e) { ExtC r; for (int i = 0; i < 5; i++) r.c[i] = bb::from_monty(e.c[i]); return r; }

struct BatchOpeningH { std::vector<std::vector<uint32_t>> opened_values; std::vector<Digest> opening_proof; };
struct CommitPhaseStepH { ExtC sibling_value; std::vector<Digest> opening_proof; };
struct QueryProofH { std::vector<CommitPhaseStepH> steps; };
struct FriProofH { std::vector<Digest> commit_phase_commits; std::vector<QueryProofH> query_proofs; ExtC final_poly; uint32_t pow_witness; };
struct OpeningH {
    std::vector<std::vector<std::vector<std::vector<ExtC>>>> values;   // [round][matrix][point][column]
    FriProofH fri;
    std::vector<std::vector<BatchOpeningH>> query_openings;             // [query][round]
};
struct OpenRound { const vgpu_prover_data* pd; std::vector<std::vector<E5>> points; };

// Device time of a phase: an event pair on the context's stream, read back by vgpu_last_prove_phases — no host
// synchronisation inside the proof.
struct Phase {
    vgpu_ctx* ctx;
    static cudaEvent_t ev(vgpu_ctx* c) {
        cudaEvent_t e = nullptr;
        if (!c->event_pool.empty()) { e = c->event_pool.back(); c->event_pool.pop_back(); } else cudaEventCreate(&e);
        return e;
    }
    Phase(vgpu_ctx* c, const char* n) : ctx(c) {
        vgpu_ctx::PhaseMark m; m.name = n; m.a = ev(c); m.b = ev(c);
        cudaEventRecord(m.a, c->stream);
        c->phase_marks.push_back(m);
        idx = c->phase_marks.size() - 1;
    }
    ~Phase() { cudaEventRecord(ctx->phase_marks[idx].b, ctx->stream); }
    size_t idx;
};
// host-side stretches between device work (pointer lists, CBOR): wall-clock, reported beside the device phases
struct HostPhase {
    vgpu_ctx* ctx; const char* name; std::chrono::steady_clock::time_point t0;
    HostPhase(vgpu_ctx* c, const char* n) : ctx(c), name(n), t0(std::chrono::steady_clock::now()) {}
    ~HostPhase() { ctx->host_phases.push_back({name, std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count()}); }
};
void phases_reset(vgpu_ctx* ctx) {
    for (auto& m : ctx->phase_marks) { ctx->event_pool.push_back(m.a); ctx->event_pool.push_back(m.b); }
    ctx->phase_marks.clear();
    ctx->phases.clear();
    ctx->host_phases.clear();
}

// An ext5 vector over the rows of one LDE height (reduced openings, inverse denominators, FRI layers): limb-major.
// Split proof: a vector of a height whose matrices are row shards holds this rank's run [begin, begin + count) only.
struct RowVec {
    uint32_t* d = nullptr;
    uint64_t n = 0, begin = 0, count = 0;      // whole length; stored run (limb stride = count)
    bool shard() const { return count != n; }
    const uint32_t* at(const vgpu_ctx* ctx, int limb, uint64_t i) const {   // null: another rank reports this element (see VgTree::node)
        if (!shard()) return ctx->comm_rank == 0 || !vg_sharded(ctx) ? d + (uint64_t)limb * count + i : nullptr;
        return i >= begin && i < begin + count ? d + (uint64_t)limb * count + (i - begin) : nullptr;
    }
};
int32_t rowvec_alloc(vgpu_ctx* ctx, uint64_t n, RowVec* v) {
    v->n = n;
    if (vg_split_rows(ctx, n)) { v->count = n / ctx->comm_size; v->begin = v->count * ctx->comm_rank; } else { v->count = n; v->begin = 0; }
    return vg_alloc(ctx, (void**)&v->d, 5 * v->count * 4);
}

struct FriLayer { RowVec values; VgTree tree; };

struct PointKey {
    uint32_t log_H; uint32_t c[5];
    bool operator<(const PointKey& o) const { if (log_H != o.log_H) return log_H < o.log_H; return std::memcmp(c, o.c, 20) < 0; }
};

int log2u(uint64_t n) { int l = 0; while ((1ull << l) < n) l++; return l; }

// Device buffers of open_multi_batches, released on every exit path.
struct OpenScratch {
    vgpu_ctx* ctx;
    std::map<PointKey, uint32_t*> invden;       // (height, point) -> 1/(x - z) over this rank's rows of the coset
    RowVec ro[32];
    uint32_t* d_sums = nullptr;
    RowVec current;
    std::vector<FriLayer> layers;
    explicit OpenScratch(vgpu_ctx* c) : ctx(c) {}
    void drop_invden() { for (auto& kv : invden) vg_free(ctx, kv.second); invden.clear(); vg_free(ctx, d_sums); d_sums = nullptr; }
    ~OpenScratch() {
        drop_invden();
        for (auto& r : ro) vg_free(ctx, r.d);
        vg_free(ctx, current.d);
        for (auto& L : layers) { vg_free(ctx, L.values.d); vg_tree_free(ctx, &L.tree); }
    }
};

int32_t open_multi_batches(vgpu_ctx* ctx, const std::vector<OpenRound>& rounds, vgh::Challenger& ch, OpeningH* out) {
    const E5 alpha = ch.sample_ext();
    // alpha^c table (host; the reduced-opening kernel takes its powers through the kernel parameters)
    uint32_t max_w = 1;
    for (auto& r : rounds) for (auto* m : r.pd->ldes) max_w = std::max<uint32_t>(max_w, (uint32_t)m->gw);
    std::vector<E5> apow(max_w + 1);
    { E5 a = bb::e5_one(); for (uint32_t c = 0; c <= max_w; c++) { apow[c] = a; a = bb::e5_mul(a, alpha); } }

    OpenScratch S(ctx);
    uint64_t num_reduced[32] = {0};
    // Pass 1 — enqueue, for every matrix, the inverse denominators of its points and the column sums behind p_c(z_q); nothing
    // here waits for the device.  One copy brings the sums of all matrices back.
    struct Job { const vgpu_dmat* lde; uint32_t log_H, w; const std::vector<E5>* pts; const uint32_t* dens[2]; size_t sums_at; };
    std::vector<Job> jobs;
    size_t sums_words = 0;
    for (const OpenRound& rd : rounds)
        for (size_t mi = 0; mi < rd.pd->ldes.size(); mi++) {
            Job j{};
            j.lde = rd.pd->ldes[mi]; j.log_H = (uint32_t)log2u(j.lde->gh); j.w = (uint32_t)j.lde->gw; j.pts = &rd.points[mi];
            if (j.pts->empty() || j.pts->size() > 2) VG_FAIL(ctx, "open: 1 or 2 points per matrix are supported");
            if ((j.lde->dist == VG_ROWS) != vg_split_rows(ctx, j.lde->gh)) VG_FAIL(ctx, "open: a committed matrix is not distributed as its height demands");
            j.sums_at = sums_words; sums_words += vg_eval_columns_words(j.w);
            jobs.push_back(j);
        }
    VG_TRY(vg_alloc(ctx, (void**)&S.d_sums, sums_words * 4));
    for (Job& j : jobs) {
        RowVec& R = S.ro[j.log_H];
        if (!R.d) {
            VG_TRY(rowvec_alloc(ctx, j.lde->gh, &R));
            VG_CUDA(ctx, cudaMemsetAsync(R.d, 0, 5 * R.count * 4, ctx->stream));
        }
        for (size_t q = 0; q < j.pts->size(); q++) {
            PointKey key; key.log_H = j.log_H; std::memcpy(key.c, (*j.pts)[q].c, 20);
            auto it = S.invden.find(key);
            if (it == S.invden.end()) {
                uint32_t* buf = nullptr;
                VG_TRY(vg_alloc(ctx, (void**)&buf, 5 * R.count * 4));
                it = S.invden.emplace(key, buf).first;
                VG_TRY(vg_inverse_denominators(ctx, j.log_H, (*j.pts)[q], R.begin, R.count, buf));
            }
            j.dens[q] = it->second;
        }
        VG_TRY(vg_eval_columns_enqueue(ctx, j.lde, (uint32_t)j.pts->size(), j.dens, R.count, S.d_sums + j.sums_at));
    }
    std::vector<uint32_t> sums(sums_words);
    VG_CUDA(ctx, cudaMemcpyAsync(sums.data(), S.d_sums, sums_words * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // Pass 2 — opened values on the host, then the reduced openings of every matrix (again without waiting)
    out->values.clear();
    {
        size_t ji = 0;
        std::vector<E5> apow_off(max_w);
        for (const OpenRound& rd : rounds) {
            out->values.emplace_back();
            for (size_t mi = 0; mi < rd.pd->ldes.size(); mi++, ji++) {
                const Job& j = jobs[ji];
                const uint32_t w = j.w, np = (uint32_t)j.pts->size();
                std::vector<E5> ys;
                vg_eval_columns_finish(sums.data() + j.sums_at, j.lde->gh, w, np, j.pts->data(), &ys, j.lde->dist == VG_ROWS ? vg_eval_columns_first_coset(w) : w);
                E5 sum_y[2];
                out->values.back().emplace_back();
                for (uint32_t q = 0; q < np; q++) {
                    E5 s = bb::e5_zero();
                    std::vector<ExtC> yc(w);
                    for (uint32_t c = 0; c < w; c++) { s = bb::e5_add(s, bb::e5_mul(apow[c], ys[q * w + c])); yc[c] = canon(ys[q * w + c]); }
                    sum_y[q] = s;
                    out->values.back().back().push_back(std::move(yc));
                }
                // point q's terms carry alpha^(num_reduced + q * w): the power table is shifted by the first offset
                const E5 a_off = bb::e5_pow(alpha, num_reduced[j.log_H]);
                num_reduced[j.log_H] += (uint64_t)w * np;
                for (uint32_t c = 0; c < w; c++) apow_off[c] = bb::e5_mul(a_off, apow[c]);
                VG_TRY(vg_reduced_opening_accumulate(ctx, j.lde, apow_off.data(), apow[w], np, j.dens, S.ro[j.log_H].count, sum_y, S.ro[j.log_H].d));
            }
        }
    }
    S.drop_invden();

    // ---- p3-fri prove: commit phase --------------------------------------------------------------
    // Split proof: while a layer is long enough it stays in row shards — a fold pairs neighbours (2i, 2i+1), so a rank folds its
    // own run, hashes its own leaves and sub-tree, and only the sub-roots meet; the first layer too short to split is
    // all-gathered once and folded by every rank from there on.
    int log_max = 31;
    while (log_max >= 0 && !S.ro[log_max].d) log_max--;
    if (log_max < LOG_BLOWUP) VG_FAIL(ctx, "open: nothing to open");
    S.current = S.ro[log_max];
    S.ro[log_max] = RowVec();
    for (int lfh = log_max - 1; lfh >= LOG_BLOWUP; lfh--) {
        S.layers.emplace_back();
        FriLayer& L = S.layers.back();
        L.values = S.current; S.current = RowVec();
        const RowVec& cur = L.values;
        const uint64_t npairs = cur.n / 2;
        Digest root;
        VG_TRY(vg_fri_layer_commit(ctx, cur.d, cur.count, npairs, cur.shard(), &L.tree, root.data()));
        ch.observe_digest_canonical(root.data());
        out->fri.commit_phase_commits.push_back(root);
        E5 beta = ch.sample_ext();
        RowVec next;
        VG_TRY(rowvec_alloc(ctx, npairs, &next));
        S.current = next;
        const RowVec& add = S.ro[lfh];
        if (add.d && add.shard() != next.shard()) VG_FAIL(ctx, "open: reduced openings and FRI layer of height 2^%d are distributed differently", lfh);
        const uint64_t i0 = cur.shard() ? cur.begin / 2 : 0, cnt = cur.count / 2;      // outputs folded here
        VG_TRY(vg_fri_fold(ctx, cur.d, cur.count, cur.n, i0, cnt, beta, add.d ? add.d - add.begin : nullptr, add.count, next.d - next.begin, next.count));
        if (cur.shard() && !next.shard()) {   // every rank folded its run into the whole-length buffer: complete it
            VG_TRY(vg_comm_group_begin(ctx));
            for (int l = 0; l < 5; l++) VG_TRY(vg_comm_allgather_inplace(ctx, next.d + (uint64_t)l * next.count, cnt));
            VG_TRY(vg_comm_group_end(ctx));
        }
        if (S.ro[lfh].d) { vg_free(ctx, S.ro[lfh].d); S.ro[lfh] = RowVec(); }
    }
    {
        const RowVec& cur = S.current;
        if (cur.shard()) VG_FAIL(ctx, "open: the final FRI layer is still distributed");
        std::vector<uint32_t> fin(5 * cur.n);
        VG_CUDA(ctx, cudaMemcpyAsync(fin.data(), cur.d, fin.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
        VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        E5 f0;
        for (int l = 0; l < 5; l++) f0.c[l] = fin[l * cur.n];
        for (uint64_t i = 1; i < cur.n; i++)
            for (int l = 0; l < 5; l++)
                if (fin[l * cur.n + i] != f0.c[l]) VG_FAIL(ctx, "FRI: final layer is not constant (the committed functions are not low degree)");
        out->fri.final_poly = canon(f0);
    }
    {
        HostPhase hp(ctx, "proof-of-work grind (device search + host check)");
        uint32_t wm = 0;
        VG_TRY(vg_pow_grind(ctx, ch, POW_BITS, &wm));
        out->fri.pow_witness = bb::from_monty(wm);
    }
    std::vector<uint64_t> indices;
    for (int q = 0; q < NUM_QUERIES; q++) indices.push_back(ch.sample_bits(log_max));

    // ---- query phase: one gather for every word the proof needs (split proof: every rank reports the words it holds) ------
    HostPhase hq(ctx, "host+device: query phase (pointer list, gather, answers)");
    // every query asks for the same NUMBER of words (its index only selects which): the 40 pointer lists, and later the 40 answers,
    // are filled by the host threads in parallel
    size_t per_query = 0;
    for (auto& L : S.layers) per_query += 5 + 8 * (L.tree.layer_ptr.size() - 1);
    for (const OpenRound& rd : rounds) { for (auto* m : rd.pd->ldes) per_query += m->w; per_query += 8 * (size_t)log2u(rd.pd->max_height); }
    const long nq = (long)indices.size();
    std::vector<const uint32_t*> ptrs(per_query * (size_t)nq);
#pragma omp parallel for schedule(static) num_threads(8)
    for (long qi = 0; qi < nq; qi++) {
        const uint64_t index = indices[qi];
        const uint32_t** o = ptrs.data() + per_query * (size_t)qi;
        auto push_digest = [&](const uint32_t* d) { for (int k = 0; k < 8; k++) *o++ = d ? d + k : nullptr; };
        for (size_t i = 0; i < S.layers.size(); i++) {
            const FriLayer& L = S.layers[i];
            uint64_t index_i = index >> i, sib = index_i ^ 1, pair = index_i >> 1;
            for (int l = 0; l < 5; l++) *o++ = L.values.at(ctx, l, sib);
            for (size_t lvl = 0; lvl + 1 < L.tree.layer_ptr.size(); lvl++) push_digest(L.tree.node(ctx, lvl, (pair >> lvl) ^ 1));
        }
        for (const OpenRound& rd : rounds) {
            int lg = log2u(rd.pd->max_height);
            uint64_t bidx = index >> (log_max - lg);
            for (auto* m : rd.pd->ldes) {
                const uint64_t row = bidx >> (lg - log2u(m->gh));
                const bool mine = m->dist == VG_ROWS ? (row >= m->row0 && row < m->row0 + m->h) : (ctx->comm_rank == 0 || !vg_sharded(ctx));
                for (uint64_t c = 0; c < m->w; c++) *o++ = mine ? m->d + c * m->col_stride + (row - m->row0) : nullptr;
            }
            for (int lvl = 0; lvl < lg; lvl++) push_digest(rd.pd->tree.node(ctx, lvl, (bidx >> lvl) ^ 1));
        }
    }
    std::vector<uint32_t> words;
    VG_TRY(vg_gather_words(ctx, ptrs, &words));
    out->fri.query_proofs.assign((size_t)nq, QueryProofH());
    out->query_openings.assign((size_t)nq, std::vector<BatchOpeningH>());
#pragma omp parallel for schedule(static) num_threads(8)
    for (long qi = 0; qi < nq; qi++) {
        size_t pos = per_query * (size_t)qi;
        auto take_digest = [&]() { Digest d; for (int k = 0; k < 8; k++) d[k] = words[pos++]; return d; };
        QueryProofH& qp = out->fri.query_proofs[qi];
        qp.steps.reserve(S.layers.size());
        for (size_t i = 0; i < S.layers.size(); i++) {
            CommitPhaseStepH st;
            for (int l = 0; l < 5; l++) st.sibling_value.c[l] = bb::from_monty(words[pos++]);
            const size_t depth = S.layers[i].tree.layer_ptr.size() - 1;
            st.opening_proof.reserve(depth);
            for (size_t lvl = 0; lvl < depth; lvl++) st.opening_proof.push_back(take_digest());
            qp.steps.push_back(std::move(st));
        }
        for (const OpenRound& rd : rounds) {
            BatchOpeningH bo;
            int lg = log2u(rd.pd->max_height);
            for (auto* m : rd.pd->ldes) {
                std::vector<uint32_t> row(m->w);
                for (uint64_t c = 0; c < m->w; c++) row[c] = bb::from_monty(words[pos++]);
                bo.opened_values.push_back(std::move(row));
            }
            for (int lvl = 0; lvl < lg; lvl++) bo.opening_proof.push_back(take_digest());
            out->query_openings[qi].push_back(std::move(bo));
        }
    }
    return 0;
}

// ---- CBOR (serde/ciborium image of MachineProof; machine/src/proof.rs:13-44) -------------------------
struct Cbor {
    // append-only byte buffer with a raw cursor: the ~150 k field elements of a proof are 12-byte stores, not push_backs
    std::vector<uint8_t> b; size_t n = 0;
    uint8_t* room(size_t k) { if (n + k > b.size()) b.resize(std::max(2 * b.size(), n + k + (1u << 20))); return b.data() + n; }
    void finish() { b.resize(n); }
    void head(uint8_t major, uint64_t v) {
        uint8_t* o = room(9);
        const uint8_t m = (uint8_t)(major << 5);
        if (v < 24) { o[0] = m | (uint8_t)v; n += 1; }
        else if (v <= 0xff) { o[0] = m | 24; o[1] = (uint8_t)v; n += 2; }
        else if (v <= 0xffff) { o[0] = m | 25; o[1] = (uint8_t)(v >> 8); o[2] = (uint8_t)v; n += 3; }
        else if (v <= 0xffffffffull) { o[0] = m | 26; for (int i = 0; i < 4; i++) o[1 + i] = (uint8_t)(v >> (24 - 8 * i)); n += 5; }
        else { o[0] = m | 27; for (int i = 0; i < 8; i++) o[1 + i] = (uint8_t)(v >> (56 - 8 * i)); n += 9; }
    }
    void key(const char* s) { size_t k = std::strlen(s); head(3, k); std::memcpy(room(k), s, k); n += k; }
    void map(uint64_t k) { head(5, k); }
    void arr(uint64_t k) { head(4, k); }
    // BabyBear { value } holds the Montgomery word: {"value": u32}
    void felt(uint32_t canonical) {
        static const uint8_t pre[7] = {0xa1, 0x65, 'v', 'a', 'l', 'u', 'e'};
        const uint32_t v = bb::to_monty(canonical);
        uint8_t* o = room(12);
        std::memcpy(o, pre, 7);
        if (v < 24) { o[7] = (uint8_t)v; n += 8; }
        else if (v <= 0xff) { o[7] = 24; o[8] = (uint8_t)v; n += 9; }
        else if (v <= 0xffff) { o[7] = 25; o[8] = (uint8_t)(v >> 8); o[9] = (uint8_t)v; n += 10; }
        else { o[7] = 26; o[8] = (uint8_t)(v >> 24); o[9] = (uint8_t)(v >> 16); o[10] = (uint8_t)(v >> 8); o[11] = (uint8_t)v; n += 12; }
    }
    void ext(const ExtC& e) { map(1); key("value"); arr(5); for (int i = 0; i < 5; i++) felt(e.c[i]); }
    void digest(const Digest& d) { arr(8); for (int i = 0; i < 8; i++) felt(d[i]); }
    void digests(const std::vector<Digest>& v) { arr(v.size()); for (auto& d : v) digest(d); }
    void exts(const std::vector<ExtC>& v) { arr(v.size()); for (auto& e : v) ext(e); }
};

// TwoAdicFriPcsProof { fri_proof, query_openings } as serde/ciborium writes it
void write_opening_proof(Cbor& w, const OpeningH& op) {
    // (encoding the per-query parts on several host threads into buffers of their own was measured: 0.88 ms against 0.64 ms serial)
    w.map(2);
    w.key("fri_proof"); w.map(4);
    w.key("commit_phase_commits"); w.digests(op.fri.commit_phase_commits);
    w.key("query_proofs"); w.arr(op.fri.query_proofs.size());
    for (auto& q : op.fri.query_proofs) {
        w.map(1); w.key("commit_phase_openings"); w.arr(q.steps.size());
        for (auto& s : q.steps) { w.map(2); w.key("sibling_value"); w.ext(s.sibling_value); w.key("opening_proof"); w.digests(s.opening_proof); }
    }
    w.key("final_poly"); w.ext(op.fri.final_poly);
    w.key("pow_witness"); w.felt(op.fri.pow_witness);
    w.key("query_openings"); w.arr(op.query_openings.size());
    for (auto& q : op.query_openings) {
        w.arr(q.size());
        for (auto& bo : q) {
            w.map(2);
            w.key("opened_values"); w.arr(bo.opened_values.size());
            for (auto& row : bo.opened_values) { w.arr(row.size()); for (uint32_t x : row) w.felt(x); }
            w.key("opening_proof"); w.digests(bo.opening_proof);
        }
    }
}

vgh::Poseidon16* poseidon_of(vgpu_ctx* ctx) {
    if (!ctx->poseidon) {
        auto* p = new vgh::Poseidon16();
        p->set(ctx->poseidon_rc, ctx->poseidon_has_mds ? ctx->poseidon_mds : nullptr);
        ctx->poseidon = p;
    }
    return (vgh::Poseidon16*)ctx->poseidon;
}

struct PdGuard { vgpu_prover_data* p = nullptr; ~PdGuard() { if (p) vgpu_prover_data_free(p); } };
struct BufGuard { vgpu_ctx* ctx; void* p = nullptr; explicit BufGuard(vgpu_ctx* c) : ctx(c) {} ~BufGuard() { vg_free(ctx, p); } };
struct MatGuard { std::vector<vgpu_dmat*> v; ~MatGuard() { for (auto* m : v) vgpu_dmat_free(m); } };

}  // namespace

void vg_host_state_free(vgpu_ctx* ctx) {
    delete (vgh::Challenger*)ctx->challenger; ctx->challenger = nullptr;
    delete (vgh::Poseidon16*)ctx->poseidon; ctx->poseidon = nullptr;
}

extern "C" {

int32_t vgpu_challenger_reset(vgpu_ctx* ctx) {
    if (!ctx->challenger_set) VG_FAIL(ctx, "challenger: vgpu_set_challenger has not been called");
    delete (vgh::Poseidon16*)ctx->poseidon; ctx->poseidon = nullptr;
    delete (vgh::Challenger*)ctx->challenger;
    auto* c = new vgh::Challenger();
    c->perm = poseidon_of(ctx);
    ctx->challenger = c;
    return 0;
}
int32_t vgpu_challenger_observe(vgpu_ctx* ctx, const uint32_t* values, uint32_t n) {
    if (!ctx->challenger) VG_TRY(vgpu_challenger_reset(ctx));
    for (uint32_t i = 0; i < n; i++) ((vgh::Challenger*)ctx->challenger)->observe(bb::to_monty(values[i] % bb::P));
    return 0;
}
int32_t vgpu_challenger_sample_ext(vgpu_ctx* ctx, uint32_t out[5]) {
    if (!ctx->challenger) VG_TRY(vgpu_challenger_reset(ctx));
    E5 e = ((vgh::Challenger*)ctx->challenger)->sample_ext();
    for (int i = 0; i < 5; i++) out[i] = bb::from_monty(e.c[i]);
    return 0;
}

int32_t vgpu_open(vgpu_ctx* ctx, const vgpu_prover_data* const* rounds, uint32_t n_rounds, const uint32_t* n_points, const uint32_t* points,
                  uint8_t** out_cbor, uint64_t* out_len) {
    if (!rounds || !n_points || !points || !out_cbor || !out_len) VG_FAIL(ctx, "open: null argument");
    VG_TRY(vg_enter(ctx));
    if (!ctx->challenger) VG_TRY(vgpu_challenger_reset(ctx));
    std::vector<OpenRound> rds(n_rounds);
    size_t mi = 0, pi = 0;
    for (uint32_t r = 0; r < n_rounds; r++) {
        if (!rounds[r]) VG_FAIL(ctx, "open: round %u has no prover data", r);
        rds[r].pd = rounds[r];
        for (size_t m = 0; m < rounds[r]->ldes.size(); m++, mi++) {
            std::vector<E5> pts;
            for (uint32_t q = 0; q < n_points[mi]; q++, pi++) {
                E5 z; for (int l = 0; l < 5; l++) z.c[l] = bb::to_monty(points[5 * pi + l] % bb::P);
                pts.push_back(z);
            }
            rds[r].points.push_back(std::move(pts));
        }
    }
    OpeningH op;
    VG_TRY(open_multi_batches(ctx, rds, *(vgh::Challenger*)ctx->challenger, &op));
    Cbor w;
    w.arr(2);
    w.arr(op.values.size());
    for (auto& round : op.values) {
        w.arr(round.size());
        for (auto& mat : round) { w.arr(mat.size()); for (auto& at_point : mat) w.exts(at_point); }
    }
    write_opening_proof(w, op);
    w.finish();
    uint8_t* buf = (uint8_t*)std::malloc(w.b.size());
    if (!buf) VG_FAIL(ctx, "out of host memory");
    std::memcpy(buf, w.b.data(), w.b.size());
    *out_cbor = buf; *out_len = w.b.size();
    return 0;
}

int32_t vgpu_prove_device(vgpu_ctx* ctx, const vgpu_dmat* const main[VGPU_NUM_CHIPS], const vgpu_dmat* const prep[2],
                          uint8_t** proof_out, uint64_t* proof_len) {
    if (!ctx->challenger_set) VG_FAIL(ctx, "prove: vgpu_set_challenger has not been called");
    if (!proof_out || !proof_len) VG_FAIL(ctx, "prove: null output");
    VG_TRY(vg_enter(ctx));
    if (!ctx->in_host_prove) phases_reset(ctx);
    const vgpu_chip_desc* chips[VGPU_NUM_CHIPS];
    int log_degrees[VGPU_NUM_CHIPS];
    for (int i = 0; i < VGPU_NUM_CHIPS; i++) {
        chips[i] = vgpu_basic_machine_chip(i);
        if (!main[i] || main[i]->gw != chips[i]->width) VG_FAIL(ctx, "prove: chip %d trace has width %llu, expected %u", i, main[i] ? (unsigned long long)main[i]->gw : 0ull, chips[i]->width);
        log_degrees[i] = log2u(main[i]->gh);
        if ((1ull << log_degrees[i]) != main[i]->gh) VG_FAIL(ctx, "prove: chip %d trace height is not a power of two", i);
    }
    if (vg_sharded(ctx)) {
        // Split proof: room in the symmetric heap for everything the three commits put there — the row shards of every tall
        // chip's main / permutation / quotient LDEs and the column buffers of the rows -> columns hand-over (counted as if never
        // released: first-fit then always finds a run) — reserved ONCE, before the first shard is live.
        std::vector<std::pair<uint64_t, uint64_t>> dm, dq, dc, dpre;
        for (int i = 0; i < VGPU_NUM_CHIPS; i++) {
            const uint64_t h = main[i]->gh;
            dm.push_back({h, chips[i]->width}); dq.push_back({h, 5ull * (chips[i]->n_interactions + 1)}); dc.push_back({h, 10});
            if (chips[i]->preprocessed_width) dpre.push_back({h, chips[i]->preprocessed_width});
        }
        const size_t need = vg_commit_symm_need(ctx, dm) + vg_commit_symm_need(ctx, dq) + vg_commit_symm_need(ctx, dc) + vg_commit_symm_need(ctx, dpre);
        if (need) VG_TRY(vg_symm_reserve(ctx, need));
    }
    vgh::Challenger ch;
    { delete (vgh::Poseidon16*)ctx->poseidon; ctx->poseidon = nullptr; }
    ch.perm = poseidon_of(ctx);

    PdGuard prep_pd, main_pd, perm_pd, quot_pd;
    uint32_t digest[8];
    {   // preprocessed commit (derive:299-311)
        Phase ph(ctx, "commit preprocessed");
        VG_TRY(vgpu_commit_batches(ctx, prep, 2, nullptr, digest, &prep_pd.p));
        ch.observe_digest_canonical(digest);
    }
    Digest main_commit, perm_commit, quot_commit;
    {   // main commit (313-332)
        Phase ph(ctx, "commit main");
        VG_TRY(vgpu_commit_batches(ctx, main, VGPU_NUM_CHIPS, nullptr, main_commit.data(), &main_pd.p));
        ch.observe_digest_canonical(main_commit.data());
    }
    uint32_t perm_challenges[15];
    for (int i = 0; i < 3; i++) { E5 e = ch.sample_ext(); for (int l = 0; l < 5; l++) perm_challenges[5 * i + l] = bb::from_monty(e.c[l]); }
    uint32_t cumsum[VGPU_NUM_CHIPS][5];
    {   // permutation traces (339-358)
        MatGuard perms;
        {
            Phase ph(ctx, "permutation traces");
            // the cumulative sums of all chips come back with ONE copy: per chip the per-rank sums of its running sum
            const uint32_t slots = vg_perm_totals_ranks(ctx);
            BufGuard tot(ctx);
            VG_TRY(vg_alloc(ctx, (void**)&tot.p, (size_t)VGPU_NUM_CHIPS * slots * 5 * 4));
            uint32_t nt[VGPU_NUM_CHIPS];
            for (int i = 0; i < VGPU_NUM_CHIPS; i++) {
                const vgpu_dmat* p = i == 1 ? prep[0] : i == 12 ? prep[1] : nullptr;
                vgpu_dmat* pm = nullptr;
                VG_TRY(vg_perm_trace_enqueue(ctx, chips[i], main[i], p, perm_challenges, &pm, (uint32_t*)tot.p + (size_t)i * slots * 5, &nt[i]));
                perms.v.push_back(pm);
            }
            std::vector<uint32_t> ht((size_t)VGPU_NUM_CHIPS * slots * 5);
            VG_CUDA(ctx, cudaMemcpyAsync(ht.data(), tot.p, ht.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
            VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            for (int i = 0; i < VGPU_NUM_CHIPS; i++)
                for (int l = 0; l < 5; l++) {
                    uint32_t a = 0;
                    for (uint32_t r = 0; r < nt[i]; r++) a = bb::add(a, ht[((size_t)i * slots + r) * 5 + l]);
                    cumsum[i][l] = bb::from_monty(a);
                }
        }
        Phase ph(ctx, "commit permutation");
        VG_TRY(vgpu_commit_batches(ctx, perms.v.data(), VGPU_NUM_CHIPS, nullptr, perm_commit.data(), &perm_pd.p));
        ch.observe_digest_canonical(perm_commit.data());
    }
    E5 alpha = ch.sample_ext();
    uint32_t alpha_c[5];
    for (int l = 0; l < 5; l++) alpha_c[l] = bb::from_monty(alpha.c[l]);
    {   // quotients (246-270, 362-374)
        MatGuard quots;
        {
            Phase ph(ctx, "quotient");
            int prep_idx = 0;
            for (int i = 0; i < VGPU_NUM_CHIPS; i++) {
                const vgpu_dmat* plde = chips[i]->preprocessed_width ? prep_pd.p->ldes[prep_idx++] : nullptr;
                vgpu_dmat* q = nullptr;
                VG_TRY(vgpu_quotient(ctx, chips[i], (uint32_t)log_degrees[i], plde, main_pd.p->ldes[i], perm_pd.p->ldes[i], cumsum[i], perm_challenges, alpha_c, &q));
                quots.v.push_back(q);
            }
        }
        Phase ph(ctx, "commit quotient");
        uint32_t shifts[VGPU_NUM_CHIPS];
        for (int i = 0; i < VGPU_NUM_CHIPS; i++) shifts[i] = (uint32_t)(((uint64_t)bb::GEN_CANON * bb::GEN_CANON) % bb::P);   // coset_shift^(2^log_quotient_degree)
        VG_TRY(vgpu_commit_batches(ctx, quots.v.data(), VGPU_NUM_CHIPS, shifts, quot_commit.data(), &quot_pd.p));
        ch.observe_digest_canonical(quot_commit.data());
    }
    E5 zeta = ch.sample_ext();
    // openings (379-392): main & perm at [zeta, zeta*g_i], quotient at [zeta^2]; preprocessed is NOT opened
    std::vector<OpenRound> rounds(3);
    rounds[0].pd = main_pd.p; rounds[1].pd = perm_pd.p; rounds[2].pd = quot_pd.p;
    for (int i = 0; i < VGPU_NUM_CHIPS; i++) {
        E5 zg = bb::e5_mul_base(zeta, bb::two_adic_generator_monty(log_degrees[i]));
        rounds[0].points.push_back({zeta, zg});
        rounds[1].points.push_back({zeta, zg});
        rounds[2].points.push_back({bb::e5_sqr(zeta)});
    }
    OpeningH op;
    {
        Phase ph(ctx, "open (evaluate + FRI)");
        VG_TRY(open_multi_batches(ctx, rounds, ch, &op));
    }
    // MachineProof -> CBOR
    HostPhase hc(ctx, "host: MachineProof -> CBOR");
    Cbor w;
    w.b.resize(4u << 20);
    w.map(3);
    w.key("commitments"); w.map(3);
    w.key("main_trace"); w.digest(main_commit);
    w.key("perm_trace"); w.digest(perm_commit);
    w.key("quotient_chunks"); w.digest(quot_commit);
    w.key("opening_proof"); write_opening_proof(w, op);
    w.key("chip_proofs"); w.arr(VGPU_NUM_CHIPS);
    const std::vector<ExtC> none;
    for (int i = 0; i < VGPU_NUM_CHIPS; i++) {
        w.map(3);
        w.key("log_degree"); w.head(0, (uint64_t)log_degrees[i]);
        w.key("opened_values"); w.map(7);
        w.key("preprocessed_local"); w.exts(none);
        w.key("preprocessed_next"); w.exts(none);
        w.key("trace_local"); w.exts(op.values[0][i][0]);
        w.key("trace_next"); w.exts(op.values[0][i][1]);
        w.key("permutation_local"); w.exts(op.values[1][i][0]);
        w.key("permutation_next"); w.exts(op.values[1][i][1]);
        w.key("quotient_chunks"); w.exts(op.values[2][i][0]);
        w.key("cumulative_sum");
        ExtC cs; for (int l = 0; l < 5; l++) cs.c[l] = cumsum[i][l];
        w.ext(cs);
    }
    w.finish();
    uint8_t* buf = (uint8_t*)std::malloc(w.b.size());
    if (!buf) VG_FAIL(ctx, "out of host memory");
    std::memcpy(buf, w.b.data(), w.b.size());
    *proof_out = buf; *proof_len = w.b.size();
    return 0;
}

int32_t vgpu_prove(vgpu_ctx* ctx, const vgpu_matrix main[VGPU_NUM_CHIPS], const vgpu_matrix prep[2], int32_t repr,
                   uint8_t** proof_out, uint64_t* proof_len) {
    // Host-buffer entry: the H2D copies run on a copy stream in the order the commits consume the traces (preprocessed,
    // then main traces tallest first — the Merkle tree hashes the tallest LDEs first and extends the shorter matrices
    // only when a tree layer of their height is reached), each transpose is deferred to the matrix's first use, so
    // copies of later matrices overlap the LDE / Keccak kernels of earlier ones.
    // Split proof: of a trace tall enough to be split a rank uploads ITS run of rows only (1 / comm_size of the bytes).
    VG_TRY(vg_enter(ctx));
    MatGuard dm, dp;
    phases_reset(ctx);
    auto t0 = std::chrono::steady_clock::now();
    dm.v.assign(VGPU_NUM_CHIPS, nullptr); dp.v.assign(2, nullptr);
    auto alloc_for = [&](const vgpu_matrix& hm, vgpu_dmat** out) {
        return vg_split_rows(ctx, 2 * hm.height) ? vg_dmat_alloc_dist(ctx, VG_ROWS, hm.height, hm.width, false, out) : vg_dmat_alloc(ctx, hm.height, hm.width, out);
    };
    auto begin_upload = [&](const vgpu_matrix& hm, vgpu_dmat* m) { return vg_upload_begin(ctx, hm.data + m->row0 * hm.width, m->h, m->w, repr, m); };
    for (int i = 0; i < 2; i++) VG_TRY(alloc_for(prep[i], &dp.v[i]));
    for (int i = 0; i < VGPU_NUM_CHIPS; i++) VG_TRY(alloc_for(main[i], &dm.v[i]));
    for (int i = 0; i < 2; i++) VG_TRY(begin_upload(prep[i], dp.v[i]));
    std::vector<int> order(VGPU_NUM_CHIPS);
    for (int i = 0; i < VGPU_NUM_CHIPS; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return main[a].height > main[b].height; });
    for (int i : order) VG_TRY(begin_upload(main[i], dm.v[i]));
    float up = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    ctx->phases.push_back({"upload traces (H2D enqueue; copies overlap the commits)", up});
    // traces in pageable memory are copied by the context's staging threads from here on (staging.cu)
    struct StagerGuard { vgpu_ctx* c; ~StagerGuard() { vg_stager_finish(c); } } sg{ctx};
    VG_TRY(vg_stager_start(ctx));
    ctx->in_host_prove = true;
    int32_t rc = vgpu_prove_device(ctx, dm.v.data(), dp.v.data(), proof_out, proof_len);
    ctx->in_host_prove = false;
    if (rc == 0) rc = vg_stager_finish(ctx);
    return rc;
}

void vgpu_free_bytes(uint8_t* p) { std::free(p); }

uint32_t vgpu_last_prove_phases(const vgpu_ctx* cctx, const char** names, float* ms, uint32_t cap) {
    vgpu_ctx* ctx = const_cast<vgpu_ctx*>(cctx);
    cudaStreamSynchronize(ctx->stream);
    for (auto& m : ctx->phase_marks) {          // drain the event pairs of the last prove into (name, milliseconds)
        float e = 0;
        if (cudaEventElapsedTime(&e, m.a, m.b) != cudaSuccess) { cudaGetLastError(); e = -1.f; }
        ctx->phases.push_back({m.name, e});
        ctx->event_pool.push_back(m.a); ctx->event_pool.push_back(m.b);
    }
    ctx->phase_marks.clear();
    for (auto& hp : ctx->host_phases) ctx->phases.push_back(hp);
    ctx->host_phases.clear();
    uint32_t n = (uint32_t)ctx->phases.size();
    for (uint32_t i = 0; i < n && i < cap; i++) { names[i] = ctx->phases[i].first; ms[i] = ctx->phases[i].second; }
    return n;
}

}  // extern "C"
