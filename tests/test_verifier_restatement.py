"""`Machine::verify` read a third time, in plain Python: the in-repo body (derive/src/lib.rs:492-650 — dimensions, transcript
order, the three rounds handed to the PCS, the cumulative-sum check) written from the Rust text, `verify_constraints`
(machine/src/verify.rs:11-107) through the transcription of tests/test_quotient_restatement.py, and underneath it the
Plonky3-side verifier as published — `TwoAdicFriPcs::verify_multi_batches`, `FieldMerkleTreeMmcs::verify_batch` for
mixed heights, p3-fri's `verify_challenges` / `verify_query` — on the hash, sponge and field helpers of
tests/test_pcs_restatement.py (own Keccak-f in Python integers, Poseidon on frequencies, the duplex sponge as a state machine).

The reference's proving tests assert exactly one thing about a proof: `machine.verify(&config, &proof)` accepts it, also after a
CBOR round trip (basic/tests/test_prover.rs:458-469).  The oracle's verifier and the product's vgpu_verify are C++ texts by one
hand; this file is the independent text for that acceptance check: it ACCEPTS the oracle's proof bytes (which the GPU proofs
equal byte for byte) and REJECTS tampered ones at the stage the C++ verifiers name.  No GPU."""
import cbor2
import numpy as np
import pytest

from test_pcs_restatement import P, R, SWITCH, Duplex, Tree, brev, commit_ldes, compress, e_pow, hash_words, two_adic_generator
from test_perm_trace_restatement import e_add, e_inv, e_mul, e_sub
from test_quotient_restatement import verify_constraints_py

RINV = pow(R, P - 2, P)
LOG_BLOWUP, NUM_QUERIES, POW_BITS, NUM_CHIPS = 1, 40, 8, 14          # FriConfig of basic/src/bin/valida.rs:385-390
ZERO5, ONE5 = [0] * 5, [1, 0, 0, 0, 0]


class Reject(Exception):
    pass


def felt(d):            # BabyBear { value: Montgomery u32 }
    v = d["value"]
    if not (isinstance(v, int) and 0 <= v < P):
        raise Reject("malformed")
    return v * RINV % P


def ext(d):
    if len(d["value"]) != 5:
        raise Reject("malformed")
    return [felt(x) for x in d["value"]]


def digest(l):
    if len(l) != 8:
        raise Reject("malformed")
    return [felt(x) for x in l]


def verify_batch(commit, heights, index, rows, path):
    """FieldMerkleTreeMmcs::verify_batch: the rows of the tallest matrices are the leaf, a sibling per level, and whenever the walk
    reaches the height of further matrices their rows are hashed and compressed in (matrices of one height in commit order)."""
    order = sorted(range(len(heights)), key=lambda i: -heights[i])              # stable
    max_h = heights[order[0]]
    if len(path) != max_h.bit_length() - 1:
        raise Reject("input_merkle")
    pos = 0

    def group(h):
        nonlocal pos
        words = []
        while pos < len(order) and heights[order[pos]] == h:
            words += rows[order[pos]]
            pos += 1
        return hash_words(words)

    node = group(max_h)
    h = max_h
    for sib in path:
        node = compress(sib, node) if index & 1 else compress(node, sib)
        index >>= 1
        h >>= 1
        if pos < len(order) and heights[order[pos]] == h:
            node = compress(node, group(h))
    return pos == len(order) and node == commit


def verify_multi_batches(rounds, pf, ch):
    """rounds: [(commit, [height], [[point]], [[values at point]])]; pf: the decoded TwoAdicFriPcsProof."""
    alpha = ch.sample_ext()
    fri = pf["fri_proof"]
    commits = [digest(c) for c in fri["commit_phase_commits"]]
    betas = []
    for c in commits:                                               # p3-fri verify_challenges
        ch.observe_digest(c)
        betas.append(ch.sample_ext())
    if len(fri["query_proofs"]) != NUM_QUERIES or len(pf["query_openings"]) != NUM_QUERIES:
        raise Reject("shape")
    ch.observe(felt(fri["pow_witness"]))                           # check_witness: observe, then the low bits of a sample must vanish
    if ch.sample_bits(POW_BITS) != 0:
        raise Reject("pow")
    log_max = len(commits) + LOG_BLOWUP
    for _, heights, _, _ in rounds:
        if any((2 * h).bit_length() - 1 > log_max for h in heights):
            raise Reject("shape")
    indices = [ch.sample_bits(log_max) for _ in range(NUM_QUERIES)]
    final_poly = ext(fri["final_poly"])
    g = SWITCH["generator"]
    for q, index in enumerate(indices):
        ro = {}
        apow = {}
        openings = pf["query_openings"][q]
        if len(openings) != len(rounds):
            raise Reject("shape")
        for (commit, heights, points, values), bo in zip(rounds, openings):
            rows = [[felt(x) for x in row] for row in bo["opened_values"]]
            if len(rows) != len(heights):
                raise Reject("shape")
            lde_heights = [h << LOG_BLOWUP for h in heights]
            lg_round = max(lde_heights).bit_length() - 1
            if not verify_batch(commit, lde_heights, index >> (log_max - lg_round), rows, [digest(d) for d in bo["opening_proof"]]):
                raise Reject("input_merkle")
            for mi, H in enumerate(lde_heights):
                lh = H.bit_length() - 1
                rev = brev(index >> (log_max - lh), lh)
                x = g * pow(two_adic_generator(lh), rev, P) % P
                for z, at_z in zip(points[mi], values[mi]):
                    if len(at_z) != len(rows[mi]):
                        raise Reject("shape")
                    den = e_sub([x, 0, 0, 0, 0], z)
                    if den == ZERO5:
                        raise Reject("shape")
                    dinv = e_inv(den)
                    for px, pz in zip(rows[mi], at_z):              # (p(x) - p(z)) / (x - z), one alpha power per (point, column)
                        quot = e_mul(e_sub([px, 0, 0, 0, 0], pz), dinv)
                        ro[lh] = e_add(ro.get(lh, ZERO5), e_mul(apow.get(lh, ONE5), quot))
                        apow[lh] = e_mul(apow.get(lh, ONE5), alpha)
        # p3-fri verify_query: fold down the layers, each pair opened in that layer's tree
        steps = fri["query_proofs"][q]["commit_phase_openings"]
        if len(steps) != len(commits):
            raise Reject("shape")
        folded = ZERO5
        x = pow(two_adic_generator(log_max), brev(index, log_max), P)
        for si, lfh in enumerate(range(log_max - 1, LOG_BLOWUP - 1, -1)):
            folded = e_add(folded, ro.get(lfh + 1, ZERO5))
            sib_slot, pair = (index ^ 1) & 1, index >> 1
            evals = [folded, folded]
            evals[sib_slot] = ext(steps[si]["sibling_value"])
            if not verify_batch(commits[si], [1 << lfh], pair, [evals[0] + evals[1]], [digest(d) for d in steps[si]["opening_proof"]]):
                raise Reject("fri_merkle")
            xs = [x, x]
            xs[sib_slot] = (P - x) % P                              # the pair sits at x and -x
            # the line through (xs[0], evals[0]), (xs[1], evals[1]) at beta
            slope = [v * pow((xs[1] - xs[0]) % P, P - 2, P) % P for v in e_sub(evals[1], evals[0])]
            folded = e_add(evals[0], e_mul(e_sub(betas[si], [xs[0], 0, 0, 0, 0]), slope))
            index = pair
            x = x * x % P
        if ro.get(LOG_BLOWUP, ZERO5) != ZERO5:                      # one-row traces: their reduced opening is exactly zero
            raise Reject("fri_final")
        if folded != final_poly:
            raise Reject("fri_final")


CHIP_WIDTHS = [51, 1, 14, 16, 16, 18, 14, 28, 45, 14, 79, 7, 2, 6]
N_INTERACTIONS = None


def machine_verify_py(proof_bytes, preprocessed, rc, n_interactions):
    try:
        d = cbor2.loads(proof_bytes)
        cps = d["chip_proofs"]
        if len(cps) != NUM_CHIPS:
            raise Reject("shape")
        log_degrees = [cp["log_degree"] for cp in cps]
        if any(not isinstance(l, int) or l < 0 or l > 26 for l in log_degrees):
            raise Reject("shape")
        if (1 << log_degrees[1]) != len(preprocessed[0]) or (1 << log_degrees[12]) != len(preprocessed[1]):
            raise Reject("shape")                                  # program ROM and range table have the height of their preprocessed columns
        ov = [{k: [ext(e) for e in v] for k, v in cp["opened_values"].items()} for cp in cps]
        if any(o["preprocessed_local"] or o["preprocessed_next"] for o in ov):
            raise Reject("shape")                                  # the reference does not open the preprocessed commitment (derive:379-392)
        cumsums = [ext(cp["cumulative_sum"]) for cp in cps]
        main_c, perm_c, quot_c = (digest(d["commitments"][k]) for k in ("main_trace", "perm_trace", "quotient_chunks"))
    except (KeyError, TypeError, ValueError, cbor2.CBORDecodeError):
        raise Reject("malformed")
    ch = Duplex(rc)
    ldes, _ = commit_ldes([[[int(v) for v in row] for row in m] for m in preprocessed])
    ch.observe_digest(Tree(ldes).root())                           # derive/src/lib.rs:585-598
    ch.observe_digest(main_c)
    perm_challenges = [ch.sample_ext() for _ in range(3)]
    ch.observe_digest(perm_c)
    alpha = ch.sample_ext()
    ch.observe_digest(quot_c)
    zeta = ch.sample_ext()
    heights = [1 << l for l in log_degrees]
    zeta_next = [e_mul(zeta, [two_adic_generator(l), 0, 0, 0, 0]) for l in log_degrees]
    zeta_sq = e_mul(zeta, zeta)                                     # zeta.exp_power_of_2(log_quotient_degree = 1)
    for i, o in enumerate(ov):                                      # widths the dimensions of derive:533-556 imply for the opened rows
        if len(o["trace_local"]) != CHIP_WIDTHS[i] or len(o["trace_next"]) != CHIP_WIDTHS[i] or len(o["quotient_chunks"]) != 10:
            raise Reject("shape")
        if len(o["permutation_local"]) != 5 * (n_interactions[i] + 1) or len(o["permutation_next"]) != 5 * (n_interactions[i] + 1):
            raise Reject("shape")
    rounds = [
        (main_c, heights, [[zeta, zn] for zn in zeta_next], [[o["trace_local"], o["trace_next"]] for o in ov]),
        (perm_c, heights, [[zeta, zn] for zn in zeta_next], [[o["permutation_local"], o["permutation_next"]] for o in ov]),
        (quot_c, heights, [[zeta_sq]] * NUM_CHIPS, [[o["quotient_chunks"]] for o in ov]),
    ]
    verify_multi_batches(rounds, d["opening_proof"], ch)
    ch15 = [v for e in perm_challenges for v in e]
    for i, o in enumerate(ov):
        folded, rhs = verify_constraints_py(i, log_degrees[i], o["trace_local"], o["trace_next"], o["permutation_local"], o["permutation_next"],
                                            o["quotient_chunks"], cumsums[i], zeta, alpha, ch15)
        if folded != rhs:
            raise Reject("constraints chip %d" % i)
    total = ZERO5
    for c in cumsums:
        total = e_add(total, c)
    if total != ZERO5:
        raise Reject("cumulative_sum")
    return True


@pytest.fixture(scope="module")
def fib3(built, oracle):
    import valida_b200 as vb
    from test_quotient_restatement import CHIPS

    t = vb.run_program(vb.fib_program(3), initial_fp=0x1000)
    proof = oracle.prove(t.main, t.preprocessed, debug_checks=False).cbor()
    return t, proof, [len(CHIPS[i]) for i in range(14)]


def test_the_python_verifier_accepts_the_oracles_proof(fib3, oracle):
    t, proof, n_inter = fib3
    assert oracle.verify(proof, t.preprocessed) == 0
    assert machine_verify_py(proof, t.preprocessed, [int(x) for x in oracle.rc480], n_inter)
    # ... also after a CBOR round trip through another encoder (basic/tests/test_prover.rs:462-469)
    assert machine_verify_py(cbor2.dumps(cbor2.loads(proof)), t.preprocessed, [int(x) for x in oracle.rc480], n_inter)


def _mutate(proof, path, fn):
    d = cbor2.loads(proof)
    node = d
    for k in path[:-1]:
        node = node[k]
    node[path[-1]] = fn(node[path[-1]])
    return cbor2.dumps(d)


def _bump(v):
    return (v + 1) % P


# the oracle verifier's return codes (oracle/pcs.h:239-303, oracle/machine.h:340-406, oracle/machine_api.inc:100); the product's
# vgpu_verify names the same stages VGPU_REJECT_* (include/valida_b200.h) and is compared with the oracle's in tests/test_gpu_verify.py
STAGE_OF_CODE = {-1000: "malformed", -1: "shape", -10: "shape", -2: "pow", -3: "input_merkle", -4: "fri_merkle", -5: "fri_final", -20: "cumulative_sum"}


@pytest.mark.parametrize("what, path", [
    ("an opened trace value at zeta", ["chip_proofs", 0, "opened_values", "trace_local", 7, "value", 0, "value"]),
    ("an opened quotient chunk", ["chip_proofs", 2, "opened_values", "quotient_chunks", 1, "value", 3, "value"]),
    ("a row opened by a query", ["opening_proof", "query_openings", 3, 0, "opened_values", 0, 5, "value"]),
    ("a sibling digest of a query path", ["opening_proof", "query_openings", 0, 1, "opening_proof", 2, 4, "value"]),
    ("a FRI sibling", ["opening_proof", "fri_proof", "query_proofs", 5, "commit_phase_openings", 1, "sibling_value", "value", 2, "value"]),
    ("the final polynomial", ["opening_proof", "fri_proof", "final_poly", "value", 0, "value"]),
    ("the proof-of-work witness", ["opening_proof", "fri_proof", "pow_witness", "value"]),
    ("a cumulative sum", ["chip_proofs", 3, "cumulative_sum", "value", 0, "value"]),
    ("the main commitment", ["commitments", "main_trace", 0, "value"]),
])
def test_the_python_verifier_rejects_tampering_like_the_cpp_verifiers(fib3, oracle, what, path):
    t, proof, n_inter = fib3
    bad = _mutate(proof, path, _bump)
    code = oracle.verify(bad, t.preprocessed)
    assert code != 0, what
    with pytest.raises(Reject) as e:
        machine_verify_py(bad, t.preprocessed, [int(x) for x in oracle.rc480], n_inter)
    want = STAGE_OF_CODE.get(code, "constraints chip %d" % (-100 - code))
    assert str(e.value) == want, (what, code, str(e.value))          # the same first failed check as the C++ verifier


def test_a_proof_that_shrinks_a_preprocessed_chip_is_a_shape_error(fib3, oracle):
    t, proof, n_inter = fib3
    bad = _mutate(proof, ["chip_proofs", 12, "log_degree"], lambda v: v - 1)
    assert oracle.verify(bad, t.preprocessed) == -1
    with pytest.raises(Reject, match="shape"):
        machine_verify_py(bad, t.preprocessed, [int(x) for x in oracle.rc480], n_inter)


@pytest.mark.parametrize("workload", ["static_data", "config5"])
def test_the_python_verifier_accepts_proofs_of_the_other_workloads(built, oracle, workload):
    # static data (prove_static_data: the static rows open the memory trace) and the multi-chip program (sub, lt family, and / or / xor
    # active: other trace heights, so other groups of matrices per tree layer and other reduced-opening heights)
    import programs
    import valida_b200 as vb
    from test_quotient_restatement import CHIPS

    if workload == "static_data":
        prog, cells = programs.static_data_program()
        t = vb.run_program(prog, initial_fp=0x1000, static_data=cells)
    else:
        t = vb.run_program(programs.config5_program(6), initial_fp=0x1000)
    proof = oracle.prove(t.main, t.preprocessed, debug_checks=False).cbor()
    assert machine_verify_py(proof, t.preprocessed, [int(x) for x in oracle.rc480], [len(CHIPS[i]) for i in range(14)])
