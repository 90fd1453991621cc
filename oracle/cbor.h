// ORACLE — TEST INFRASTRUCTURE ONLY (see field.h header).
// CBOR image of MachineProof as `ciborium::into_writer(&proof, ..)` would emit it
// (basic/src/bin/valida.rs:425-426; structs machine/src/proof.rs:13-44 + p3-fri proof structs):
// serde structs -> definite-length maps keyed by field name, Vec / [T;N] -> definite-length arrays,
// BabyBear -> {"value": u32 in Montgomery form (x * 2^32 mod p)}, BinomialExtensionField -> {"value": [5 x BabyBear]}.
// [P3-UNVERIFIED; SURVEY App. A items 1, 20] — byte-level parity with the Rust CLI is unpinned.
#pragma once
#include "machine.h"
#include <string>
#include <stdexcept>

namespace orc {

struct CborEnc {
    std::vector<uint8_t> out;
    void head(uint8_t major, uint64_t v) {
        uint8_t m = (uint8_t)(major << 5);
        if (v < 24) out.push_back(m | (uint8_t)v);
        else if (v < 256) { out.push_back(m | 24); out.push_back((uint8_t)v); }
        else if (v < 65536) { out.push_back(m | 25); out.push_back((uint8_t)(v >> 8)); out.push_back((uint8_t)v); }
        else if (v < (1ull << 32)) { out.push_back(m | 26); for (int i = 3; i >= 0; i--) out.push_back((uint8_t)(v >> (8 * i))); }
        else { out.push_back(m | 27); for (int i = 7; i >= 0; i--) out.push_back((uint8_t)(v >> (8 * i))); }
    }
    void uint(uint64_t v) { head(0, v); }
    void array(uint64_t n) { head(4, n); }
    void map(uint64_t n) { head(5, n); }
    void text(const char* s) { size_t n = std::strlen(s); head(3, n); out.insert(out.end(), s, s + n); }
};
struct CborDec {
    const uint8_t* p; const uint8_t* end;
    uint64_t head(uint8_t major) {
        if (p >= end) throw std::runtime_error("cbor: truncated");
        uint8_t b = *p++;
        if ((b >> 5) != major) throw std::runtime_error("cbor: unexpected major type");
        uint8_t ai = b & 31;
        if (ai < 24) return ai;
        int nb = ai == 24 ? 1 : ai == 25 ? 2 : ai == 26 ? 4 : ai == 27 ? 8 : -1;
        if (nb < 0 || end - p < nb) throw std::runtime_error("cbor: bad length");
        uint64_t v = 0;
        for (int i = 0; i < nb; i++) v = (v << 8) | *p++;
        return v;
    }
    uint64_t uint() { return head(0); }
    uint64_t array() { return head(4); }
    void map(uint64_t n) { if (head(5) != n) throw std::runtime_error("cbor: map size"); }
    void key(const char* s) {
        uint64_t n = head(3);
        if (n != std::strlen(s) || (uint64_t)(end - p) < n || std::memcmp(p, s, n) != 0) throw std::runtime_error(std::string("cbor: expected key ") + s);
        p += n;
    }
};

constexpr uint32_t MONTY_R = (uint32_t)((1ull << 32) % P);
static inline uint32_t to_monty(uint32_t x) { return mul(x, MONTY_R); }
static inline uint32_t from_monty(uint32_t x) { static const uint32_t rinv = inv(MONTY_R); return mul(x, rinv); }

static inline void enc_bb(CborEnc& e, uint32_t x) { e.map(1); e.text("value"); e.uint(to_monty(x)); }
static inline uint32_t dec_bb(CborDec& d) { d.map(1); d.key("value"); uint64_t v = d.uint(); if (v >= P) throw std::runtime_error("cbor: non-canonical field element"); return from_monty((uint32_t)v); }
static inline void enc_ext(CborEnc& e, const Ext5& x) { e.map(1); e.text("value"); e.array(5); for (int i = 0; i < 5; i++) enc_bb(e, x.c[i]); }
static inline Ext5 dec_ext(CborDec& d) { Ext5 x; d.map(1); d.key("value"); if (d.array() != 5) throw std::runtime_error("cbor: ext len"); for (int i = 0; i < 5; i++) x.c[i] = dec_bb(d); return x; }
static inline void enc_digest(CborEnc& e, const Digest& g) { e.array(8); for (int i = 0; i < 8; i++) enc_bb(e, g[i]); }
static inline Digest dec_digest(CborDec& d) { Digest g; if (d.array() != 8) throw std::runtime_error("cbor: digest len"); for (int i = 0; i < 8; i++) g[i] = dec_bb(d); return g; }
static inline void enc_digests(CborEnc& e, const std::vector<Digest>& v) { e.array(v.size()); for (auto& g : v) enc_digest(e, g); }
static inline std::vector<Digest> dec_digests(CborDec& d) { uint64_t n = d.array(); std::vector<Digest> v; for (uint64_t i = 0; i < n; i++) v.push_back(dec_digest(d)); return v; }
static inline void enc_exts(CborEnc& e, const std::vector<Ext5>& v) { e.array(v.size()); for (auto& x : v) enc_ext(e, x); }
static inline std::vector<Ext5> dec_exts(CborDec& d) { uint64_t n = d.array(); std::vector<Ext5> v; for (uint64_t i = 0; i < n; i++) v.push_back(dec_ext(d)); return v; }

static inline void enc_opening_proof(CborEnc& e, const PcsProof& op) {
    e.map(2);
    const FriProof& f = op.fri;
    e.text("fri_proof"); e.map(4);
    e.text("commit_phase_commits"); enc_digests(e, f.commit_phase_commits);
    e.text("query_proofs"); e.array(f.query_proofs.size());
    for (auto& q : f.query_proofs) {
        e.map(1); e.text("commit_phase_openings"); e.array(q.steps.size());
        for (auto& s : q.steps) { e.map(2); e.text("sibling_value"); enc_ext(e, s.sibling_value); e.text("opening_proof"); enc_digests(e, s.opening_proof); }
    }
    e.text("final_poly"); enc_ext(e, f.final_poly);
    e.text("pow_witness"); enc_bb(e, f.pow_witness);
    e.text("query_openings"); e.array(op.query_openings.size());
    for (auto& q : op.query_openings) {
        e.array(q.size());
        for (auto& b : q) {
            e.map(2);
            e.text("opened_values"); e.array(b.opened_values.size());
            for (auto& row : b.opened_values) { e.array(row.size()); for (uint32_t x : row) enc_bb(e, x); }
            e.text("opening_proof"); enc_digests(e, b.opening_proof);
        }
    }
}
// the Rust tuple (OpenedValues, TwoAdicFriPcsProof) returned by open_multi_batches: a 2-element array
static inline std::vector<uint8_t> encode_opening(const OpenedValues& values, const PcsProof& proof) {
    CborEnc e;
    e.array(2);
    e.array(values.size());
    for (auto& round : values) {
        e.array(round.size());
        for (auto& mat : round) { e.array(mat.size()); for (auto& at_point : mat) enc_exts(e, at_point); }
    }
    enc_opening_proof(e, proof);
    return e.out;
}

static inline std::vector<uint8_t> encode_proof(const MachineProof& pr) {
    CborEnc e;
    e.map(3);
    e.text("commitments"); e.map(3);
    e.text("main_trace"); enc_digest(e, pr.main_trace);
    e.text("perm_trace"); enc_digest(e, pr.perm_trace);
    e.text("quotient_chunks"); enc_digest(e, pr.quotient_chunks);
    e.text("opening_proof"); enc_opening_proof(e, pr.opening_proof);
    e.text("chip_proofs"); e.array(pr.chip_proofs.size());
    for (auto& c : pr.chip_proofs) {
        e.map(3);
        e.text("log_degree"); e.uint(c.log_degree);
        e.text("opened_values"); e.map(7);
        e.text("preprocessed_local"); enc_exts(e, c.opened.preprocessed_local);
        e.text("preprocessed_next"); enc_exts(e, c.opened.preprocessed_next);
        e.text("trace_local"); enc_exts(e, c.opened.trace_local);
        e.text("trace_next"); enc_exts(e, c.opened.trace_next);
        e.text("permutation_local"); enc_exts(e, c.opened.permutation_local);
        e.text("permutation_next"); enc_exts(e, c.opened.permutation_next);
        e.text("quotient_chunks"); enc_exts(e, c.opened.quotient_chunks);
        e.text("cumulative_sum"); enc_ext(e, c.cumulative_sum);
    }
    return e.out;
}

static inline MachineProof decode_proof(const uint8_t* bytes, size_t len) {
    CborDec d{bytes, bytes + len};
    MachineProof pr;
    d.map(3);
    d.key("commitments"); d.map(3);
    d.key("main_trace"); pr.main_trace = dec_digest(d);
    d.key("perm_trace"); pr.perm_trace = dec_digest(d);
    d.key("quotient_chunks"); pr.quotient_chunks = dec_digest(d);
    d.key("opening_proof"); d.map(2);
    {
        FriProof& f = pr.opening_proof.fri;
        d.key("fri_proof"); d.map(4);
        d.key("commit_phase_commits"); f.commit_phase_commits = dec_digests(d);
        d.key("query_proofs"); uint64_t nq = d.array();
        for (uint64_t i = 0; i < nq; i++) {
            QueryProof q;
            d.map(1); d.key("commit_phase_openings"); uint64_t ns = d.array();
            for (uint64_t k = 0; k < ns; k++) { CommitPhaseStep s; d.map(2); d.key("sibling_value"); s.sibling_value = dec_ext(d); d.key("opening_proof"); s.opening_proof = dec_digests(d); q.steps.push_back(std::move(s)); }
            f.query_proofs.push_back(std::move(q));
        }
        d.key("final_poly"); f.final_poly = dec_ext(d);
        d.key("pow_witness"); f.pow_witness = dec_bb(d);
        d.key("query_openings"); uint64_t nqo = d.array();
        for (uint64_t i = 0; i < nqo; i++) {
            uint64_t nr = d.array();
            std::vector<BatchOpening> rounds;
            for (uint64_t r = 0; r < nr; r++) {
                BatchOpening b;
                d.map(2);
                d.key("opened_values"); uint64_t nm = d.array();
                for (uint64_t m = 0; m < nm; m++) { uint64_t w = d.array(); std::vector<uint32_t> row; for (uint64_t c = 0; c < w; c++) row.push_back(dec_bb(d)); b.opened_values.push_back(std::move(row)); }
                d.key("opening_proof"); b.opening_proof = dec_digests(d);
                rounds.push_back(std::move(b));
            }
            pr.opening_proof.query_openings.push_back(std::move(rounds));
        }
    }
    d.key("chip_proofs"); uint64_t nc = d.array();
    for (uint64_t i = 0; i < nc; i++) {
        ChipProof c;
        d.map(3);
        d.key("log_degree"); c.log_degree = d.uint();
        d.key("opened_values"); d.map(7);
        d.key("preprocessed_local"); c.opened.preprocessed_local = dec_exts(d);
        d.key("preprocessed_next"); c.opened.preprocessed_next = dec_exts(d);
        d.key("trace_local"); c.opened.trace_local = dec_exts(d);
        d.key("trace_next"); c.opened.trace_next = dec_exts(d);
        d.key("permutation_local"); c.opened.permutation_local = dec_exts(d);
        d.key("permutation_next"); c.opened.permutation_next = dec_exts(d);
        d.key("quotient_chunks"); c.opened.quotient_chunks = dec_exts(d);
        d.key("cumulative_sum"); c.cumulative_sum = dec_ext(d);
        pr.chip_proofs.push_back(std::move(c));
    }
    if (d.p != d.end) throw std::runtime_error("cbor: trailing bytes");
    return pr;
}

}  // namespace orc
