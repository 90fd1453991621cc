// ORACLE — TEST INFRASTRUCTURE ONLY (see field.h header).
// TwoAdicFriPcs<TwoAdicFriPcsConfig<Val, Challenge, Challenger, Dft, ValMmcs, ChallengeMmcs>> with
// FriConfig{log_blowup:1, num_queries:40, proof_of_work_bits:8} — basic/src/bin/valida.rs:384-395.
// Methods restated: the ones Machine::prove/verify call through UnivariatePcsWithLde
// (machine/src/config.rs:17-22; derive/src/lib.rs:309,330,355,372,391-392,620-633):
// commit_batches, commit_shifted_batches (per-matrix shift slice — fork-specific), get_ldes,
// coset_shift, log_blowup, open_multi_batches, verify_multi_batches; plus p3-fri's prover/verifier
// and fold_even_odd.   [P3-UNVERIFIED; SURVEY App. A items 5, 9, 14, 15, 17]
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "field.h"
#include "ntt.h"
#include "merkle.h"
#include "poseidon.h"

namespace orc {

struct FriConfig { int log_blowup = 1; int num_queries = 40; int pow_bits = 8; };

struct PcsData { MerkleTree tree; };  // leaves: bit-reversed coset LDEs, caller's order

struct CommitPhaseStep { Ext5 sibling_value; std::vector<Digest> opening_proof; };
struct QueryProof { std::vector<CommitPhaseStep> steps; };
struct FriProof {
    std::vector<Digest> commit_phase_commits;
    std::vector<QueryProof> query_proofs;
    Ext5 final_poly;
    uint32_t pow_witness;
};
struct PcsProof {
    FriProof fri;
    std::vector<std::vector<BatchOpening>> query_openings;  // [query][round]
};
// [round][matrix][point][column]
using OpenedValues = std::vector<std::vector<std::vector<std::vector<Ext5>>>>;
using Points = std::vector<std::vector<Ext5>>;  // [matrix][point]

// ExtensionMmcs: an ext-field matrix of width 2 is committed as its base flattening (width 10).
static inline Matrix flatten_ext_pairs(const std::vector<Ext5>& v) {
    Matrix m(v.size() / 2, 10);
#pragma omp parallel for schedule(static) if (v.size() > 8192)
    for (long i = 0; i < (long)v.size(); i++) std::memcpy(&m.v[(size_t)i * 5], v[i].c, 20);
    return m;
}

// p3-fri fold_even_odd: input bit-reversed evaluations; pairs (2i, 2i+1) are (p(x), p(-x)).
static inline std::vector<Ext5> fold_even_odd(const std::vector<Ext5>& poly, const Ext5& beta) {
    size_t half = poly.size() / 2;
    int log_half = log2_strict(half);
    uint32_t g_inv = inv(two_adic_generator(log_half + 1));
    uint32_t one_half = inv(2);
    Ext5 half_beta = beta * one_half;
    std::vector<Ext5> powers(half);
    {
        std::vector<uint32_t> gp = geometric(1, g_inv, half);
#pragma omp parallel for schedule(static) if (half > 4096)
        for (long i = 0; i < (long)half; i++) powers[reverse_bits_len((uint32_t)i, log_half)] = half_beta * gp[i];
    }
    std::vector<Ext5> out(half);
#pragma omp parallel for schedule(static) if (half > 4096)
    for (long i = 0; i < (long)half; i++) {
        const Ext5& lo = poly[2 * i];
        const Ext5& hi = poly[2 * i + 1];
        out[i] = (powers[i] + one_half) * lo + (-powers[i] + one_half) * hi;
    }
    return out;
}

struct Pcs {
    FriConfig fri;
    uint32_t coset_shift() const { return GEN; }
    int log_blowup() const { return fri.log_blowup; }

    PcsData commit_shifted_batches(const std::vector<Matrix>& polys, const std::vector<uint32_t>& coset_shifts) const {
        std::vector<Matrix> ldes;
        for (size_t i = 0; i < polys.size(); i++) {
            uint32_t shift = mul(GEN, inv(coset_shifts[i]));
            Matrix l = coset_lde_batch(polys[i], fri.log_blowup, shift);
            bit_reverse_rows(l);
            ldes.push_back(std::move(l));
        }
        PcsData d;
        d.tree = merkle_commit(std::move(ldes));
        return d;
    }
    PcsData commit_batches(const std::vector<Matrix>& polys) const {
        return commit_shifted_batches(polys, std::vector<uint32_t>(polys.size(), 1));
    }
    // get_ldes: natural-order row r of LDE i is stored row reverse_bits(r).
    static inline const uint32_t* lde_row(const Matrix& bitrev_lde, size_t r) {
        return bitrev_lde.row(reverse_bits_len((uint32_t)r, log2_strict(bitrev_lde.height())));
    }

    // p(z) for every column, from the first h rows of the bit-reversed LDE (= evaluations over g*H).
    std::vector<Ext5> eval_at(const Matrix& lde, const Ext5& z) const {
        size_t h = lde.height() >> fri.log_blowup, w = lde.width;
        int lg = log2_strict(h);
        uint32_t s = GEN, om = two_adic_generator(lg);
        std::vector<uint32_t> xs = geometric(s, om, h);
        std::vector<Ext5> den(h);
#pragma omp parallel for schedule(static) if (h > 4096)
        for (long i = 0; i < (long)h; i++) den[i] = z - xs[i];
        std::vector<Ext5> dinv = batch_inverse(den);
        std::vector<Ext5> acc(w, Ext5::zero());
#pragma omp parallel if (h > 4096)
        {
            std::vector<Lazy5> loc(w);
            int pending = 0;
#pragma omp for schedule(static) nowait
            for (long i = 0; i < (long)h; i++) {
                Ext5 wgt = dinv[i] * xs[i];
                const uint32_t* row = lde.row(reverse_bits_len((uint32_t)i, lg));
                for (size_t c = 0; c < w; c++) loc[c].mad(wgt, row[c]);
                if (++pending == 4) { for (size_t c = 0; c < w; c++) loc[c].fold(); pending = 0; }
            }
#pragma omp critical
            for (size_t c = 0; c < w; c++) acc[c] += loc[c].value();
        }
        uint32_t sn = exp_pow2(s, lg);
        Ext5 scale = (ext_exp_pow2(z, lg) - sn) * inv(mul((uint32_t)(h % P), sn));
        for (size_t c = 0; c < w; c++) acc[c] = acc[c] * scale;
        return acc;
    }

    struct Round { const PcsData* data; Points points; };

    std::pair<OpenedValues, PcsProof> open_multi_batches(const std::vector<Round>& rounds, Challenger& ch) const {
        Ext5 alpha = ch.sample_ext();
        double T_eval = 0, T_inv = 0, T_red = 0, T_xs = 0; auto NOW = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        double T_start = NOW();
        OpenedValues all;
        std::vector<std::vector<Ext5>> ro(32);
        size_t num_reduced[32] = {0};
        for (const Round& rd : rounds) {
            all.emplace_back();
            const auto& mats = rd.data->tree.leaves;
            for (size_t mi = 0; mi < mats.size(); mi++) {
                const Matrix& mat = mats[mi];
                size_t H = mat.height(), w = mat.width;
                int lh = log2_strict(H);
                if (ro[lh].empty()) ro[lh].assign(H, Ext5::zero());
                all.back().emplace_back();
                // x for storage row i: g * omega^{bitrev(i)}
                double tq0 = NOW();
                std::vector<uint32_t> xs(H);
                { std::vector<uint32_t> nat = geometric(GEN, two_adic_generator(lh), H);
#pragma omp parallel for schedule(static) if (H > 8192)
                  for (long i = 0; i < (long)H; i++) xs[i] = nat[reverse_bits_len((uint32_t)i, lh)]; }
                std::vector<Ext5> apow(w);
                { Ext5 a = Ext5::one(); for (size_t c = 0; c < w; c++) { apow[c] = a; a = a * alpha; } }
                T_xs += NOW() - tq0;
                for (const Ext5& z : rd.points[mi]) {
                    double t0 = NOW();
                    std::vector<Ext5> ys = eval_at(mat, z);
                    double t1 = NOW(); T_eval += t1 - t0;
                    Ext5 alpha_pow_offset = ext_pow(alpha, num_reduced[lh]);
                    Ext5 sum_y = Ext5::zero();
                    for (size_t c = 0; c < w; c++) sum_y += apow[c] * ys[c];
                    std::vector<Ext5> den(H);
#pragma omp parallel for schedule(static) if (H > 4096)
                    for (long i = 0; i < (long)H; i++) den[i] = -z + xs[i];
                    std::vector<Ext5> dinv = batch_inverse(den);
                    double t2 = NOW(); T_inv += t2 - t1;
                    std::vector<Ext5>& r = ro[lh];
#pragma omp parallel for schedule(static) if (H * w > (1u << 14))
                    for (long i = 0; i < (long)H; i++) {
                        const uint32_t* row = mat.row(i);
                        Lazy5 lz;
                        for (size_t c = 0; c < w; c++) { lz.mad(apow[c], row[c]); if ((c & 3) == 3) lz.fold(); }
                        Ext5 red = lz.value();
                        r[i] += alpha_pow_offset * (red - sum_y) * dinv[i];
                    }
                    T_red += NOW() - t2;
                    num_reduced[lh] += w;
                    all.back().back().push_back(std::move(ys));
                }
            }
        }
        if (getenv("ORACLE_TIMING")) fprintf(stderr, "  open: xs %.3f eval %.3f inv %.3f reduce %.3f (total so far %.3f)\n", T_xs, T_eval, T_inv, T_red, NOW() - T_start);
        // ---- p3-fri prove ----
        int log_max_height = 31;
        while (log_max_height >= 0 && ro[log_max_height].empty()) log_max_height--;
        assert(log_max_height >= fri.log_blowup);
        PcsProof proof;
        std::vector<MerkleTree> layer_trees;
        std::vector<Ext5> current = ro[log_max_height];
        for (int lfh = log_max_height - 1; lfh >= fri.log_blowup; lfh--) {
            MerkleTree t = merkle_commit({flatten_ext_pairs(current)});
            Digest c = t.root();
            ch.observe_digest(c.data());
            proof.fri.commit_phase_commits.push_back(c);
            layer_trees.push_back(std::move(t));
            Ext5 beta = ch.sample_ext();
            current = fold_even_odd(current, beta);
            if (!ro[lfh].empty()) {
                const std::vector<Ext5>& add = ro[lfh];
#pragma omp parallel for schedule(static) if (current.size() > 8192)
                for (long i = 0; i < (long)current.size(); i++) current[i] += add[i];
            }
        }
        assert(current.size() == (1u << fri.log_blowup));
        for (auto& x : current) { assert(x == current[0]); (void)x; }
        proof.fri.final_poly = current[0];
        if (getenv("ORACLE_TIMING")) fprintf(stderr, "  open: fri commit phase done at %.3f\n", NOW() - T_start);
        proof.fri.pow_witness = ch.grind(fri.pow_bits);
        std::vector<size_t> query_indices;
        for (int q = 0; q < fri.num_queries; q++) query_indices.push_back(ch.sample_bits(log_max_height));
        for (size_t index : query_indices) {
            QueryProof qp;
            for (size_t i = 0; i < layer_trees.size(); i++) {
                size_t index_i = index >> i, index_pair = index_i >> 1;
                BatchOpening bo = merkle_open(layer_trees[i], index_pair);
                CommitPhaseStep st;
                std::memcpy(st.sibling_value.c, &bo.opened_values[0][((index_i ^ 1) & 1) * 5], 20);
                st.opening_proof = bo.opening_proof;
                qp.steps.push_back(std::move(st));
            }
            proof.fri.query_proofs.push_back(std::move(qp));
        }
        for (size_t index : query_indices) {
            proof.query_openings.emplace_back();
            for (const Round& rd : rounds) {
                int lg = log2_ceil(rd.data->tree.max_height());
                proof.query_openings.back().push_back(merkle_open(rd.data->tree, index >> (log_max_height - lg)));
            }
        }
        return {std::move(all), std::move(proof)};
    }

    struct VRound { Digest commit; Points points; std::vector<Dims> dims; /* unextended heights */ };

    // Returns 0 on success, a negative code naming the failing check otherwise.
    int verify_multi_batches(const std::vector<VRound>& rounds, const OpenedValues& values, const PcsProof& proof, Challenger& ch) const {
        Ext5 alpha = ch.sample_ext();
        std::vector<Ext5> betas;
        for (const Digest& c : proof.fri.commit_phase_commits) { ch.observe_digest(c.data()); betas.push_back(ch.sample_ext()); }
        if ((int)proof.fri.query_proofs.size() != fri.num_queries) return -1;
        if ((int)proof.query_openings.size() != fri.num_queries) return -1;
        if (!ch.check_witness(fri.pow_bits, proof.fri.pow_witness)) return -2;
        int log_max_height = (int)proof.fri.commit_phase_commits.size() + fri.log_blowup;
        std::vector<size_t> idx;
        for (int q = 0; q < fri.num_queries; q++) idx.push_back(ch.sample_bits(log_max_height));
        if (values.size() != rounds.size()) return -1;
        for (int q = 0; q < fri.num_queries; q++) {
            size_t index = idx[q];
            Ext5 ro[32]; Ext5 apw[32];
            for (int i = 0; i < 32; i++) { ro[i] = Ext5::zero(); apw[i] = Ext5::one(); }
            if (proof.query_openings[q].size() != rounds.size()) return -1;
            for (size_t r = 0; r < rounds.size(); r++) {
                const VRound& vr = rounds[r];
                const BatchOpening& bo = proof.query_openings[q][r];
                std::vector<Dims> ext_dims;
                size_t maxh = 0;
                for (auto& d : vr.dims) { ext_dims.push_back({d.width, d.height << fri.log_blowup}); maxh = std::max(maxh, d.height << fri.log_blowup); }
                size_t bidx = index >> (log_max_height - log2_ceil(maxh));
                if (!merkle_verify(vr.commit, ext_dims, bidx, bo.opened_values, bo.opening_proof)) return -3;
                if (values[r].size() != vr.dims.size()) return -1;
                for (size_t mi = 0; mi < vr.dims.size(); mi++) {
                    int lh = log2_strict(vr.dims[mi].height) + fri.log_blowup;
                    int bits_reduced = log_max_height - lh;
                    uint32_t rev = reverse_bits_len((uint32_t)(index >> bits_reduced), lh);
                    uint32_t x = mul(GEN, pw(two_adic_generator(lh), rev));
                    if (values[r][mi].size() != vr.points[mi].size()) return -1;
                    for (size_t pi = 0; pi < vr.points[mi].size(); pi++) {
                        const Ext5& z = vr.points[mi][pi];
                        const std::vector<Ext5>& ps_at_z = values[r][mi][pi];
                        if (ps_at_z.size() != bo.opened_values[mi].size()) return -1;
                        Ext5 dinv = ext_inv(-z + x);
                        for (size_t c = 0; c < ps_at_z.size(); c++) {
                            Ext5 quotient = (-ps_at_z[c] + bo.opened_values[mi][c]) * dinv;
                            ro[lh] += apw[lh] * quotient;
                            apw[lh] = apw[lh] * alpha;
                        }
                    }
                }
            }
            // p3-fri verify_query
            const QueryProof& qp = proof.fri.query_proofs[q];
            if (qp.steps.size() != proof.fri.commit_phase_commits.size()) return -1;
            Ext5 folded = Ext5::zero();
            uint32_t x = pw(two_adic_generator(log_max_height), reverse_bits_len((uint32_t)index, log_max_height));
            size_t si = 0;
            for (int lfh = log_max_height - 1; lfh >= fri.log_blowup; lfh--, si++) {
                folded += ro[lfh + 1];
                size_t index_sibling = index ^ 1, index_pair = index >> 1;
                Ext5 evals[2] = {folded, folded};
                evals[index_sibling % 2] = qp.steps[si].sibling_value;
                std::vector<uint32_t> rowv(10);
                std::memcpy(&rowv[0], evals[0].c, 20); std::memcpy(&rowv[5], evals[1].c, 20);
                if (!merkle_verify(proof.fri.commit_phase_commits[si], {{10, (size_t)1 << lfh}}, index_pair, {rowv}, qp.steps[si].opening_proof)) return -4;
                uint32_t xs[2] = {x, x};
                xs[index_sibling % 2] = mul(xs[index_sibling % 2], two_adic_generator(1));
                // interpolate through (xs[0],evals[0]),(xs[1],evals[1]) and evaluate at beta
                folded = evals[0] + (betas[si] - xs[0]) * (evals[1] - evals[0]) * inv(sub(xs[1], xs[0]));
                index = index_pair;
                x = mul(x, x);
            }
            // reduced openings of the height-2 LDEs (one-row traces: constant polynomials, exactly 0 when honest) are added by
            // the prover's last fold but never reached by the loop above: bind them here
            if (!ro[fri.log_blowup].is_zero()) return -5;
            if (folded != proof.fri.final_poly) return -5;
        }
        return 0;
    }
};

}  // namespace orc
