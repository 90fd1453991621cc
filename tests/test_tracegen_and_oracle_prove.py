"""Host witness generation + oracle prove/verify (CPU only): reproduces the reference's own
prove_fibonacci assertions (basic/tests/test_prover.rs:474-487) and its debug-build invariants."""
import cbor2
import numpy as np
import pytest

P = 2013265921


@pytest.fixture(scope="module")
def fib25(built):
    import valida_b200 as vb

    return vb.run_program(vb.fib_program(25), initial_fp=0x1000)


def test_fib_program_vm_state_matches_reference_test(fib25):
    t = fib25
    assert t.clock == 192                      # machine.cpu().clock == 192
    assert t.mem_ops == 401                    # mem().operations ... count() == 401
    assert t.add_ops == 105                    # add_u32().operations.len() == 105
    assert t.mem_cell(0x1000 + 4) == 75025     # Word([0, 1, 37, 17]) — 25th Fibonacci number
    assert [m.shape for m in t.main] == [(256, 51), (32, 1), (512, 14), (128, 16), (1, 16), (1024, 18), (1, 14), (1, 28), (1, 45), (1, 14), (1, 79), (1, 7), (256, 2), (1, 6)]
    assert t.preprocessed[0].shape == (32, 7) and t.preprocessed[1].shape == (256, 1)
    assert all(int(m.max()) < P for m in t.main)


def test_workload_model_cycles(built):
    # cycles = 17 + 7n, memory ops = 26 + 15n, adds = 5 + 4n
    import valida_b200 as vb

    for n in [0, 1, 2, 10, 582]:
        t = vb.run_program(vb.fib_program(n))
        assert (t.clock, t.mem_ops, t.add_ops) == (17 + 7 * n, 26 + 15 * n, 5 + 4 * n)


def fib(n):
    a, b = 0, 1
    for _ in range(n):
        a, b = b, (a + b) % (1 << 32)
    return a


def test_fib_results_wrap_mod_2_32(built):
    import valida_b200 as vb

    for n in [0, 1, 5, 47, 48, 100]:
        assert vb.run_program(vb.fib_program(n)).mem_cell(0x1004) == fib(n)


def test_oracle_proves_and_verifies_fib25(fib25, oracle):
    pr = oracle.prove(fib25.main, fib25.preprocessed, debug_checks=True)
    # debug-build invariants of the reference: every constraint vanishes on every row; cumulative sums cancel
    assert pr.constraint_failures() == [-1] * 14
    assert pr.cumulative_sum_zero()
    proof = pr.cbor()
    assert oracle.verify(proof, fib25.preprocessed) == 0
    d = cbor2.loads(proof)
    assert sorted(d.keys()) == ["chip_proofs", "commitments", "opening_proof"]
    assert len(d["chip_proofs"]) == 14
    assert [c["log_degree"] for c in d["chip_proofs"]] == [8, 5, 9, 7, 0, 10, 0, 0, 0, 0, 0, 0, 8, 0]
    fri = d["opening_proof"]["fri_proof"]
    assert len(fri["query_proofs"]) == 40 and len(fri["commit_phase_commits"]) == 10  # log2(2048) - log_blowup
    # determinism
    assert oracle.prove(fib25.main, fib25.preprocessed, debug_checks=False).cbor() == proof


def _flip_value(obj, path):
    """Add 1 (mod p) to the BabyBear leaf reached by `path` inside the decoded CBOR proof."""
    cur = obj
    for k in path[:-1]:
        cur = cur[k]
    leaf = cur[path[-1]]
    leaf["value"] = (leaf["value"] + 1) % P


def test_verifier_rejects_tampering(fib25, oracle):
    proof = oracle.prove(fib25.main, fib25.preprocessed, debug_checks=False).cbor()
    targets = [
        ["commitments", "main_trace", 0],
        ["commitments", "quotient_chunks", 7],
        ["chip_proofs", 0, "opened_values", "trace_local", 3, "value", 0],
        ["chip_proofs", 2, "opened_values", "permutation_next", 1, "value", 4],
        ["chip_proofs", 3, "opened_values", "quotient_chunks", 9, "value", 2],
        ["chip_proofs", 0, "cumulative_sum", "value", 0],
        ["opening_proof", "fri_proof", "final_poly", "value", 0],
        ["opening_proof", "fri_proof", "pow_witness"],
        ["opening_proof", "fri_proof", "query_proofs", 5, "commit_phase_openings", 2, "sibling_value", "value", 1],
        ["opening_proof", "fri_proof", "query_proofs", 0, "commit_phase_openings", 0, "opening_proof", 0, 0],
        ["opening_proof", "query_openings", 3, 1, "opened_values", 2, 0],
        ["opening_proof", "query_openings", 0, 0, "opening_proof", 4, 5],
    ]
    for path in targets:
        d = cbor2.loads(proof)
        if path[-1] == "pow_witness":
            d["opening_proof"]["fri_proof"]["pow_witness"]["value"] = (d["opening_proof"]["fri_proof"]["pow_witness"]["value"] + 1) % P
        else:
            _flip_value(d, path)
        bad = cbor2.dumps(d)
        assert oracle.verify(bad, fib25.preprocessed) != 0, path
    # a corrupted witness is caught by the debug invariants and by the verifier
    bad_main = [m.copy() for m in fib25.main]
    bad_main[3][5, 11] = (int(bad_main[3][5, 11]) + 1) % P  # an ADD output byte
    pr = oracle.prove(bad_main, fib25.preprocessed, debug_checks=True)
    assert pr.constraint_failures()[3] != -1
    assert oracle.verify(pr.cbor(), fib25.preprocessed) != 0


def test_verifier_binds_one_row_chips_and_preprocessed_heights(fib25, oracle):
    """A chip with a ONE-row trace has height-2 LDEs whose reduced openings no FRI round folds; the verifier must still
    check them (they are exactly 0 for an honest proof), or the opened values and the cumulative sum of such a chip are
    free — eight of the fourteen chips in a Fibonacci proof.  Likewise the program / range chips keep the height of
    their preprocessed columns."""
    proof = oracle.prove(fib25.main, fib25.preprocessed, debug_checks=False).cbor()
    assert fib25.main[13].shape[0] == 1 and fib25.main[6].shape[0] == 1
    for chip, which, col in ((13, "trace_local", 2), (6, "trace_next", 5), (11, "permutation_local", 0), (9, "quotient_chunks", 3)):
        d = cbor2.loads(proof)
        _flip_value(d, ["chip_proofs", chip, "opened_values", which, col, "value", 0])
        assert oracle.verify(cbor2.dumps(d), fib25.preprocessed) == -5, (chip, which)      # the opening check itself, not a later one
    for chip in (1, 12):
        d = cbor2.loads(proof)
        d["chip_proofs"][chip]["log_degree"] = 0
        assert oracle.verify(cbor2.dumps(d), fib25.preprocessed) == -1


# ---- the reference's other proving tests (basic/tests/test_prover.rs:490-625), via tests/golden/programs.json ----
import json
import os

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "programs.json")))


@pytest.mark.parametrize("name", ["left_imm_ops_program", "signed_inequality_program", "loadfp_program"])
def test_reference_test_programs_vm_state_and_oracle_proof(built, oracle, name):
    import valida_b200 as vb

    g = GOLDEN[name]
    t = vb.run_program(np.array(g["program"], dtype=np.int32), initial_fp=0x1000)
    for addr, value in g["expected_cells"]:          # the reference test's own assertions on machine.mem().cells
        assert t.mem_cell(addr) == value, (name, hex(addr))
    pr = oracle.prove(t.main, t.preprocessed, debug_checks=True)
    assert pr.constraint_failures() == [-1] * 14     # check_constraints (debug builds of the reference)
    assert pr.cumulative_sum_zero()                  # check_cumulative_sums
    assert oracle.verify(pr.cbor(), t.preprocessed) == 0


def test_config5_program_trace_satisfies_every_air(built, oracle):
    """add + sub + lt family + and/or/xor: the host witness generator's rows satisfy the restated AIRs on every row
    (the reference's debug-build check_constraints), all LogUp sums cancel, and the oracle proof verifies."""
    import valida_b200 as vb
    from programs import config5_program

    t = vb.run_program(config5_program(60), initial_fp=0x1000)
    # 60 iterations: 2 sub, 6 bitwise, 4 lt, 2 add per iteration
    assert t.main[4].shape[0] == 128 and t.main[10].shape[0] == 512 and t.main[8].shape[0] == 256 and t.main[3].shape[0] == 128
    # VM state: replay the loop in Python
    M = 0xFFFFFFFF
    x, m = 0x9E3779B9, 0x0FF055AA
    for i in range(60):
        x = (x + 1013904223) & M
        y = x ^ m
        s = x | 0x80808080
        e = (s - 12345) & M
        x ^= e
        m = ((m & y) | 0x01010101) & M
    assert t.mem_cell(0x1000 - 8) == x and t.mem_cell(0x1000 - 32) == m and t.mem_cell(0x1000 - 4) == 60
    ref = oracle.prove(t.main, t.preprocessed, debug_checks=True)
    assert ref.constraint_failures() == [-1] * 14 and ref.cumulative_sum_zero()
    assert oracle.verify(ref.cbor(), t.preprocessed) == 0
