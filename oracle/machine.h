// ORACLE — TEST INFRASTRUCTURE ONLY (see field.h header).
// Machine::prove / Machine::verify for BasicMachine, restated from the reference:
//   prove orchestration + transcript   derive/src/lib.rs:275-446 (hand copy basic/src/lib.rs:147-675)
//   generate_permutation_trace         machine/src/chip.rs:121-208
//   eval_permutation_constraints       machine/src/chip.rs:210-289, generate_rlc_elements 291-331
//   check_constraints / cumulative     machine/src/check_constraints.rs:14-93
//   quotient / quotient_values         machine/src/quotient.rs:18-238
//   decompose_and_flatten, ZerofierOnCoset   p3-uni-stark [P3-UNVERIFIED; SURVEY App. A item 16]
//   verify / verify_constraints        derive/src/lib.rs:492-650, machine/src/verify.rs:11-107
// Inputs are the 14 main traces (row-major, canonical u32) and the two preprocessed traces
// (program: 7 columns, range: 1 column) — trace generation is outside this file.
#pragma once
#include <ctime>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "field.h"
#include "air.h"
#include "pcs.h"
#include "poseidon.h"

namespace orc {

struct ExtMatrix {  // RowMajorMatrix<Challenge>
    std::vector<Ext5> v; size_t width = 0;
    size_t height() const { return width ? v.size() / width : 0; }
    Matrix flatten_to_base() const {
        Matrix m(height(), width * 5);
#pragma omp parallel for schedule(static) if (v.size() > 8192)
        for (long i = 0; i < (long)v.size(); i++) std::memcpy(&m.v[(size_t)i * 5], v[i].c, 20);
        return m;
    }
};

// generate_rlc_elements (machine/src/chip.rs:291-331): alphas_global[i] = r1^(i+1), i <= max bus index.
static inline std::vector<Ext5> rlc_alphas_global(const ChipDef& chip, const Ext5& r1) {
    uint32_t maxbus = 0;
    for (auto& it : chip.interactions) maxbus = std::max(maxbus, it.bus);
    std::vector<Ext5> a;
    Ext5 p = r1;
    for (uint32_t i = 0; i <= maxbus; i++) { a.push_back(p); p = p * r1; }
    return a;
}

static inline ExtMatrix generate_permutation_trace(const ChipDef& chip, const Matrix& main, const Matrix* prep, const Ext5 rnd[3]) {
    size_t h = main.height(), k = chip.interactions.size(), pw_ = k + 1;
    std::vector<Ext5> alphas = rlc_alphas_global(chip, rnd[1]);
    size_t maxf = 0;
    for (auto& it : chip.interactions) maxf = std::max(maxf, it.fields.size());
    std::vector<Ext5> betas(maxf);
    { Ext5 b = Ext5::one(); for (size_t j = 0; j < maxf; j++) { betas[j] = b; b = b * rnd[2]; } }
    ExtMatrix perm; perm.width = pw_; perm.v.assign(h * pw_, Ext5::zero());
    std::vector<Fp> mrow(main.width), prow(prep ? prep->width : 0);
#pragma omp parallel for schedule(static) firstprivate(mrow, prow) if (h > 1024)
    for (long n = 0; n < (long)h; n++) {
        for (size_t c = 0; c < main.width; c++) mrow[c] = Fp(main.at(n, c));
        if (prep) for (size_t c = 0; c < prep->width; c++) prow[c] = Fp(prep->at(n, c));
        for (size_t m = 0; m < k; m++) {
            const Interaction& it = chip.interactions[m];
            Ext5 rlc = Ext5::zero();
            for (size_t j = 0; j < it.fields.size(); j++) rlc += betas[j] * it.fields[j].apply<Fp>(prow.data(), mrow.data()).v;
            rlc += alphas[it.bus];
            perm.v[n * pw_ + m] = rlc;
        }
    }
    // batch_multiplicative_inverse_allowing_zero over the whole matrix (the phi column is still zero)
    perm.v = batch_inverse_allowing_zero(perm.v);
    // the row terms on all host threads, then the running sum itself (additions only) in row order, as the reference's loop
#pragma omp parallel for schedule(static) firstprivate(mrow, prow) if (h > 1024)
    for (long n = 0; n < (long)h; n++) {
        for (size_t c = 0; c < main.width; c++) mrow[c] = Fp(main.at(n, c));
        if (prep) for (size_t c = 0; c < prep->width; c++) prow[c] = Fp(prep->at(n, c));
        Ext5 term = Ext5::zero();
        for (size_t m = 0; m < k; m++) {
            const Interaction& it = chip.interactions[m];
            uint32_t mult = it.count.apply<Fp>(prow.data(), mrow.data()).v;
            Ext5 t = perm.v[n * pw_ + m] * mult;
            if (it.is_send) term += t; else term -= t;
        }
        perm.v[n * pw_ + k] = term;
    }
    Ext5 phi = Ext5::zero();
    for (size_t n = 0; n < h; n++) { phi += perm.v[n * pw_ + k]; perm.v[n * pw_ + k] = phi; }
    return perm;
}

template <class E>
static inline void eval_permutation_constraints(const ChipDef& chip, Folder<E>& b, const Ext5& cumulative_sum) {
    const Ext5* rnd = b.perm_challenges;
    size_t k = chip.interactions.size(), pw_ = b.perm_width;
    const Ext5& phi_local = b.perm_local[pw_ - 1];
    const Ext5& phi_next = b.perm_next[pw_ - 1];
    std::vector<Ext5> alphas = rlc_alphas_global(chip, rnd[1]);
    Ext5 lhs = phi_next - phi_local, rhs = Ext5::zero(), phi_0 = Ext5::zero();
    for (size_t m = 0; m < k; m++) {
        const Interaction& it = chip.interactions[m];
        Ext5 rlc = Ext5::zero();
        Ext5 beta = Ext5::one();
        for (size_t j = 0; j < it.fields.size(); j++) {
            E elem = it.fields[j].template apply<E>(b.prep_local, b.main_local);
            rlc += Lift<E>::emul(beta, elem);
            beta = beta * rnd[2];
        }
        rlc = rlc + alphas[it.bus];
        b.assert_zero_ext(rlc * b.perm_local[m] - Ext5::one());
        E mult_local = it.count.template apply<E>(b.prep_local, b.main_local);
        E mult_next = it.count.template apply<E>(b.prep_next, b.main_next);
        if (it.is_send) { phi_0 += Lift<E>::emul(b.perm_local[m], mult_local); rhs += Lift<E>::emul(b.perm_next[m], mult_next); }
        else            { phi_0 -= Lift<E>::emul(b.perm_local[m], mult_local); rhs -= Lift<E>::emul(b.perm_next[m], mult_next); }
    }
    b.when_transition().assert_eq_ext(lhs, rhs);
    b.when_first_row().assert_eq_ext(b.perm_local[pw_ - 1], phi_0);
    b.when_last_row().assert_eq_ext(b.perm_local[pw_ - 1], cumulative_sum);
}

// check_constraints (debug builds of the reference): returns -1 if every constraint vanishes on
// every row, else row * 4096 + constraint index of the first failure.
static inline long check_constraints(const ChipDef& chip, const Matrix& main, const Matrix* prep, const ExtMatrix& perm, const Ext5 rnd[3]) {
    size_t h = main.height();
    if (h == 0) return -1;
    Ext5 cumsum = perm.v[(h - 1) * perm.width + perm.width - 1];
    long bad = -1;
    for (size_t i = 0; i < h && bad < 0; i++) {
        size_t j = (i + 1) % h;
        std::vector<Fp> ml(main.width), mn(main.width), pl(prep ? prep->width : 0), pn(prep ? prep->width : 0);
        for (size_t c = 0; c < main.width; c++) { ml[c] = Fp(main.at(i, c)); mn[c] = Fp(main.at(j, c)); }
        if (prep) for (size_t c = 0; c < prep->width; c++) { pl[c] = Fp(prep->at(i, c)); pn[c] = Fp(prep->at(j, c)); }
        Folder<Fp> f;
        f.debug = true;
        f.main_local = ml.data(); f.main_next = mn.data(); f.prep_local = pl.data(); f.prep_next = pn.data();
        f.perm_local = &perm.v[i * perm.width]; f.perm_next = &perm.v[j * perm.width]; f.perm_width = perm.width;
        f.perm_challenges = rnd;
        f.is_first_row = Fp(i == 0); f.is_last_row = Fp(i == h - 1); f.is_transition = Fp(i == h - 1 ? 0 : 1);
        chip.eval_fp(f);
        eval_permutation_constraints<Fp>(chip, f, cumsum);
        if (f.first_failed >= 0) bad = (long)i * 4096 + f.first_failed;
    }
    return bad;
}

// quotient(): chunks matrix h x 10 (log_quotient_degree = 1 for every BasicMachine chip:
// machine/src/symbolic/symbolic_builder.rs:17-30 gives max(deg,3)-1 = 2 -> log2_ceil = 1).
constexpr int LOG_QUOTIENT_DEGREE = 1;

static inline Matrix quotient(const ChipDef& chip, int log_degree, const Matrix* prep_lde, const Matrix& main_lde, const Matrix& perm_lde,
                              const Ext5& cumulative_sum, const Ext5 rnd[3], const Ext5& alpha, std::vector<Ext5>* values_out = nullptr) {
    // *_lde: committed bit-reversed LDEs (2h rows); natural row r is stored at reverse_bits(r).
    const int lqd = LOG_QUOTIENT_DEGREE;
    size_t n = (size_t)1 << log_degree, qs = n << lqd;
    int log_qs = log_degree + lqd;
    uint32_t g_sub = two_adic_generator(log_degree), g_ext = two_adic_generator(log_qs);
    uint32_t subgroup_last = inv(g_sub), s = GEN;
    size_t next_step = (size_t)1 << lqd;
    uint32_t s_pow_n = exp_pow2(s, log_degree);
    // ZerofierOnCoset: Z_H(s g_ext^i) = s^n * v^(i mod rate) - 1, v = rate-th root of unity
    uint32_t zevals[2] = {sub(s_pow_n, 1), sub(mul(s_pow_n, two_adic_generator(lqd)), 1)};
    uint32_t zinv[2] = {inv(zevals[0]), inv(zevals[1])};
    std::vector<uint32_t> coset = geometric(s, g_ext, qs);
    std::vector<uint32_t> den_first(qs), den_last(qs);
#pragma omp parallel for schedule(static) if (qs > 8192)
    for (long i = 0; i < (long)qs; i++) { den_first[i] = sub(coset[i], 1); den_last[i] = sub(coset[i], subgroup_last); }  // g_h^(n-1) = g_h^-1
    std::vector<uint32_t> inv_first = batch_inverse(den_first), inv_last = batch_inverse(den_last);
    size_t w = main_lde.width, pwb = perm_lde.width, pwe = pwb / 5, wp = prep_lde ? prep_lde->width : 0;
    std::vector<Ext5> q(qs);
#pragma omp parallel for schedule(static) if (qs > 512)
    for (long i = 0; i < (long)qs; i++) {
        size_t inext = ((size_t)i + next_step) % qs;
        size_t ri = reverse_bits_len((uint32_t)i, log_qs), rn = reverse_bits_len((uint32_t)inext, log_qs);
        std::vector<Fp> ml(w), mn(w), pl(wp), pn(wp);
        std::vector<Ext5> el(pwe), en(pwe);
        for (size_t c = 0; c < w; c++) { ml[c] = Fp(main_lde.at(ri, c)); mn[c] = Fp(main_lde.at(rn, c)); }
        for (size_t c = 0; c < wp; c++) { pl[c] = Fp(prep_lde->at(ri, c)); pn[c] = Fp(prep_lde->at(rn, c)); }
        for (size_t c = 0; c < pwe; c++) { std::memcpy(el[c].c, perm_lde.row(ri) + 5 * c, 20); std::memcpy(en[c].c, perm_lde.row(rn) + 5 * c, 20); }
        Folder<Fp> f;
        f.main_local = ml.data(); f.main_next = mn.data(); f.prep_local = pl.data(); f.prep_next = pn.data();
        f.perm_local = el.data(); f.perm_next = en.data(); f.perm_width = pwe; f.perm_challenges = rnd;
        uint32_t zh = zevals[i & 1];
        f.is_first_row = Fp(mul(zh, inv_first[i]));
        f.is_last_row = Fp(mul(zh, inv_last[i]));
        f.is_transition = Fp(sub(coset[i], subgroup_last));
        f.alpha = alpha;
        chip.eval_fp(f);
        eval_permutation_constraints<Fp>(chip, f, cumulative_sum);
        q[i] = f.accumulator * zinv[i & 1];
    }
    if (values_out) *values_out = q;
    // decompose_and_flatten(q, shift = s, log_chunks = 1): even/odd halves over the coset s^2 H
    Matrix out(n, 10);
    uint32_t g_inv = inv(g_ext), one_half = inv(2);
    std::vector<uint32_t> gps = geometric(mul(inv(s), one_half), g_inv, n);
#pragma omp parallel for schedule(static) if (n > 4096)
    for (long i = 0; i < (long)n; i++) {
        Ext5 even = (q[i] + q[i + n]) * one_half;
        Ext5 odd = (q[i] - q[i + n]) * gps[i];
        std::memcpy(out.row(i), even.c, 20);
        std::memcpy(out.row(i) + 5, odd.c, 20);
    }
    return out;
}

// ---- proof structs (machine/src/proof.rs:13-44) -------------------------------------------------
struct ChipOpenedValues {
    std::vector<Ext5> preprocessed_local, preprocessed_next, trace_local, trace_next, permutation_local, permutation_next, quotient_chunks;
};
struct ChipProof { size_t log_degree; ChipOpenedValues opened; Ext5 cumulative_sum; };
struct MachineProof {
    Digest main_trace, perm_trace, quotient_chunks;
    PcsProof opening_proof;
    std::vector<ChipProof> chip_proofs;
};

// Stage dumps for stage-by-stage parity tests against the CUDA path.
struct ProveTrace {
    Digest prep_commit;
    Ext5 perm_challenges[3], alpha, zeta;
    std::vector<ExtMatrix> perm_traces;
    std::vector<Matrix> quotient_chunks;
    std::vector<long> constraint_failures;  // per chip, -1 = all vanish
    bool cumulative_sum_zero = false;
};

struct MachineInput {
    std::vector<Matrix> main_traces;   // 14, chip order
    Matrix program_prep, range_prep;   // preprocessed traces (program 7 cols, range 1 col)
};

static inline const Matrix* prep_of(const MachineInput& in, int chip) {
    if (chip == CHIP_PROGRAM) return &in.program_prep;
    if (chip == CHIP_RANGE) return &in.range_prep;
    return nullptr;
}

static inline MachineProof machine_prove(const MachineInput& in, const Poseidon16& perm16, bool debug_checks, ProveTrace* tr = nullptr) {
    const auto& cd = chips();
    Pcs pcs;
    Challenger ch(&perm16);
    // ORACLE_TIMING=1: wall time of each phase on stderr (where the CPU baseline spends its time)
    // (with OMP_WAIT_POLICY=passive the process CPU time beside it shows how much of a phase ran on one thread only)
    auto T0 = std::chrono::steady_clock::now();
    auto cpu_now = [] { timespec ts; clock_gettime(CLOCK_PROCESS_CPUTIME_ID, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
    double C0 = cpu_now();
    auto lap = [&](const char* what) { if (getenv("ORACLE_TIMING")) { auto t = std::chrono::steady_clock::now(); double c = cpu_now(); fprintf(stderr, "oracle %-18s %.3f s (cpu %.3f s)\n", what, std::chrono::duration<double>(t - T0).count(), c - C0); T0 = t; C0 = c; } };
    // preprocessed commit (derive:299-311)
    PcsData prep_data = pcs.commit_batches({in.program_prep, in.range_prep});
    Digest prep_commit = prep_data.tree.root();
    ch.observe_digest(prep_commit.data());
    // main commit (313-332)
    std::vector<int> log_degrees(NUM_CHIPS);
    for (int i = 0; i < NUM_CHIPS; i++) { assert(in.main_traces[i].width == cd[i].width); log_degrees[i] = log2_strict(in.main_traces[i].height()); }
    PcsData main_data = pcs.commit_batches(in.main_traces);
    Digest main_commit = main_data.tree.root();
    ch.observe_digest(main_commit.data());
    lap("commit main");
    Ext5 rnd[3];
    for (int i = 0; i < 3; i++) rnd[i] = ch.sample_ext();
    // permutation traces (339-358)
    std::vector<ExtMatrix> perm_traces(NUM_CHIPS);
    std::vector<Ext5> cumulative_sums(NUM_CHIPS);
    std::vector<Matrix> perm_flat(NUM_CHIPS);
    for (int i = 0; i < NUM_CHIPS; i++) {
        perm_traces[i] = generate_permutation_trace(cd[i], in.main_traces[i], prep_of(in, i), rnd);
        cumulative_sums[i] = perm_traces[i].v.back();
        perm_flat[i] = perm_traces[i].flatten_to_base();
    }
    lap("perm traces");
    PcsData perm_data = pcs.commit_batches(perm_flat);
    lap("commit perm");
    Digest perm_commit = perm_data.tree.root();
    ch.observe_digest(perm_commit.data());
    Ext5 alpha = ch.sample_ext();
    // quotients (246-270, 362-374)
    std::vector<Matrix> quotients(NUM_CHIPS);
    std::vector<long> failures(NUM_CHIPS, -1);
    int prep_idx = 0;
    for (int i = 0; i < NUM_CHIPS; i++) {
        if (debug_checks) failures[i] = check_constraints(cd[i], in.main_traces[i], prep_of(in, i), perm_traces[i], rnd);
        const Matrix* plde = cd[i].prep_width ? &prep_data.tree.leaves[prep_idx++] : nullptr;
        quotients[i] = quotient(cd[i], log_degrees[i], plde, main_data.tree.leaves[i], perm_data.tree.leaves[i], cumulative_sums[i], rnd, alpha);
    }
    lap("quotient");
    std::vector<uint32_t> coset_shifts(NUM_CHIPS, exp_pow2(pcs.coset_shift(), LOG_QUOTIENT_DEGREE));
    PcsData quot_data = pcs.commit_shifted_batches(quotients, coset_shifts);
    Digest quot_commit = quot_data.tree.root();
    ch.observe_digest(quot_commit.data());
    lap("commit quotient");
    Ext5 zeta = ch.sample_ext();
    // openings (379-392): preprocessed NOT opened (TODO in the reference)
    Points zeta_and_next(NUM_CHIPS), zeta_exp(NUM_CHIPS);
    for (int i = 0; i < NUM_CHIPS; i++) {
        uint32_t g = two_adic_generator(log_degrees[i]);
        zeta_and_next[i] = {zeta, zeta * g};
        zeta_exp[i] = {ext_exp_pow2(zeta, LOG_QUOTIENT_DEGREE)};
    }
    std::vector<Pcs::Round> rounds = {{&main_data, zeta_and_next}, {&perm_data, zeta_and_next}, {&quot_data, zeta_exp}};
    auto opened = pcs.open_multi_batches(rounds, ch);
    lap("open");
    MachineProof proof;
    proof.main_trace = main_commit; proof.perm_trace = perm_commit; proof.quotient_chunks = quot_commit;
    proof.opening_proof = std::move(opened.second);
    for (int i = 0; i < NUM_CHIPS; i++) {
        ChipProof cp;
        cp.log_degree = log_degrees[i];
        cp.opened.trace_local = opened.first[0][i][0]; cp.opened.trace_next = opened.first[0][i][1];
        cp.opened.permutation_local = opened.first[1][i][0]; cp.opened.permutation_next = opened.first[1][i][1];
        cp.opened.quotient_chunks = opened.first[2][i][0];
        cp.cumulative_sum = cumulative_sums[i];
        proof.chip_proofs.push_back(std::move(cp));
    }
    if (tr) {
        tr->prep_commit = prep_commit;
        for (int i = 0; i < 3; i++) tr->perm_challenges[i] = rnd[i];
        tr->alpha = alpha; tr->zeta = zeta;
        tr->perm_traces = perm_traces;
        tr->quotient_chunks = quotients;
        tr->constraint_failures = failures;
        Ext5 sum = Ext5::zero();
        for (auto& c : cumulative_sums) sum += c;
        tr->cumulative_sum_zero = sum.is_zero();
    }
    return proof;
}

// verify_constraints (machine/src/verify.rs:11-107); 0 = ok
static inline int verify_constraints(const ChipDef& chip, const ChipOpenedValues& ov, const Ext5& cumulative_sum, int log_degree,
                                     const Ext5& zeta, const Ext5& alpha, const Ext5 rnd[3]) {
    uint32_t g = two_adic_generator(log_degree);
    Ext5 z_h = ext_exp_pow2(zeta, log_degree) - 1u;
    Ext5 is_first_row = z_h / (zeta - 1u);
    Ext5 is_last_row = z_h / (zeta - inv(g));
    Ext5 is_transition = zeta - inv(g);
    auto unflatten = [](const std::vector<Ext5>& v) {
        std::vector<Ext5> out;
        for (size_t i = 0; i + 5 <= v.size(); i += 5) {
            Ext5 s = Ext5::zero();
            for (int k = 0; k < 5; k++) s += v[i + k] * Ext5::monomial(k);
            out.push_back(s);
        }
        return out;
    };
    if (ov.trace_local.size() != chip.width || ov.trace_next.size() != chip.width) return -10;
    size_t pw_ = chip.interactions.size() + 1;
    if (ov.permutation_local.size() != pw_ * 5 || ov.permutation_next.size() != pw_ * 5) return -10;
    if (ov.quotient_chunks.size() != 10) return -10;
    std::vector<Ext5> quotient_parts = unflatten(ov.quotient_chunks);
    std::vector<Ext5> pl = unflatten(ov.permutation_local), pn = unflatten(ov.permutation_next);
    Folder<Ext5> f;
    f.main_local = ov.trace_local.data(); f.main_next = ov.trace_next.data();
    f.prep_local = ov.preprocessed_local.data(); f.prep_next = ov.preprocessed_next.data();
    f.perm_local = pl.data(); f.perm_next = pn.data(); f.perm_width = pw_;
    f.perm_challenges = rnd;
    f.is_first_row = is_first_row; f.is_last_row = is_last_row; f.is_transition = is_transition;
    f.alpha = alpha;
    chip.eval_ext(f);
    eval_permutation_constraints<Ext5>(chip, f, cumulative_sum);
    // reverse_slice_index_bits on 2 parts is the identity
    Ext5 quot = Ext5::zero(), wgt = Ext5::one();
    for (auto& p : quotient_parts) { quot += p * wgt; wgt = wgt * zeta; }
    return f.accumulator == z_h * quot ? 0 : -11;
}

// Machine::verify (derive/src/lib.rs:492-650). 0 = accept; negative = reason (-100-chip for a chip's constraints).
static inline int machine_verify(const MachineProof& proof, const Matrix& program_prep, const Matrix& range_prep, const Poseidon16& perm16) {
    const auto& cd = chips();
    if (proof.chip_proofs.size() != (size_t)NUM_CHIPS) return -1;
    // chips with preprocessed columns (program ROM, range table) have the height of those columns
    if (((size_t)1 << proof.chip_proofs[1].log_degree) != program_prep.height() || ((size_t)1 << proof.chip_proofs[12].log_degree) != range_prep.height()) return -1;
    Pcs pcs;
    Challenger ch(&perm16);
    PcsData prep_data = pcs.commit_batches({program_prep, range_prep});
    Digest prep_commit = prep_data.tree.root();
    ch.observe_digest(prep_commit.data());
    ch.observe_digest(proof.main_trace.data());
    Ext5 rnd[3];
    for (int i = 0; i < 3; i++) rnd[i] = ch.sample_ext();
    ch.observe_digest(proof.perm_trace.data());
    Ext5 alpha = ch.sample_ext();
    ch.observe_digest(proof.quotient_chunks.data());
    Ext5 zeta = ch.sample_ext();
    std::vector<Pcs::VRound> rounds(3);
    OpenedValues values(3);
    rounds[0].commit = proof.main_trace; rounds[1].commit = proof.perm_trace; rounds[2].commit = proof.quotient_chunks;
    for (int i = 0; i < NUM_CHIPS; i++) {
        const ChipProof& cp = proof.chip_proofs[i];
        if (cp.log_degree > 26) return -1;
        size_t h = (size_t)1 << cp.log_degree;
        uint32_t g = two_adic_generator((int)cp.log_degree);
        rounds[0].dims.push_back({cd[i].width, h});
        rounds[1].dims.push_back({(cd[i].interactions.size() + 1) * 5, h});
        rounds[2].dims.push_back({10, h});
        rounds[0].points.push_back({zeta, zeta * g});
        rounds[1].points.push_back({zeta, zeta * g});
        rounds[2].points.push_back({ext_exp_pow2(zeta, LOG_QUOTIENT_DEGREE)});
        values[0].push_back({cp.opened.trace_local, cp.opened.trace_next});
        values[1].push_back({cp.opened.permutation_local, cp.opened.permutation_next});
        values[2].push_back({cp.opened.quotient_chunks});
    }
    int rc = pcs.verify_multi_batches(rounds, values, proof.opening_proof, ch);
    if (rc != 0) return rc;
    for (int i = 0; i < NUM_CHIPS; i++) {
        const ChipProof& cp = proof.chip_proofs[i];
        int r = verify_constraints(cd[i], cp.opened, cp.cumulative_sum, (int)cp.log_degree, zeta, alpha, rnd);
        if (r != 0) return -100 - i;
    }
    Ext5 sum = Ext5::zero();
    for (auto& cp : proof.chip_proofs) sum += cp.cumulative_sum;
    if (!sum.is_zero()) return -20;
    return 0;
}

}  // namespace orc
