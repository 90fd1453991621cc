"""The library's own Machine::verify (vgpu_verify) against the oracle verifier: both accept honest proofs (from
either prover), both reject the same tampered proofs, and for the same reason."""
import cbor2
import pytest

pytestmark = pytest.mark.gpu

# oracle verdicts (oracle/machine.h machine_verify, oracle/pcs.h verify_multi_batches) -> VGPU_REJECT_* codes
ORACLE_TO_VGPU = {0: 0, -1000: -1, -1: -2, -2: -3, -3: -4, -4: -5, -5: -6, -20: -7}


def product_verdict(vb, cfg, proof, prep):
    try:
        vb.verify_machine(cfg, proof, prep)
        return 0
    except vb.VerificationError as e:
        return e.verdict


def expected(oracle_code):
    return oracle_code if oracle_code <= -100 else ORACLE_TO_VGPU[oracle_code]


@pytest.fixture(scope="module")
def fib25(ctx, oracle):
    import valida_b200 as vb

    t = vb.run_program(vb.fib_program(25), initial_fp=0x1000)
    cfg = vb.StarkConfig(ctx, oracle.rc480)
    return vb, cfg, t, vb.prove_machine(cfg, t)


def test_accepts_own_and_oracle_proofs(fib25, oracle):
    vb, cfg, t, proof = fib25
    vb.verify_machine(cfg, proof, t.preprocessed)
    ref = oracle.prove(t.main, t.preprocessed, debug_checks=False).cbor()
    vb.verify_machine(cfg, ref, t.preprocessed)


def test_accepts_larger_trace(ctx, oracle):
    import valida_b200 as vb

    t = vb.run_program(vb.fib_program(582), initial_fp=0x1000)
    cfg = vb.StarkConfig(ctx, oracle.rc480)
    vb.verify_machine(cfg, vb.prove_machine(cfg, t), t.preprocessed)


def _flip(x):
    x["value"] ^= 1


TAMPERS = {
    "trace_local": lambda d: _flip(d["chip_proofs"][0]["opened_values"]["trace_local"][7]["value"][0]),
    "trace_next": lambda d: _flip(d["chip_proofs"][2]["opened_values"]["trace_next"][1]["value"][2]),
    "permutation_local": lambda d: _flip(d["chip_proofs"][3]["opened_values"]["permutation_local"][0]["value"][0]),
    "quotient_chunk": lambda d: _flip(d["chip_proofs"][0]["opened_values"]["quotient_chunks"][3]["value"][1]),
    "cumulative_sum": lambda d: _flip(d["chip_proofs"][12]["cumulative_sum"]["value"][0]),
    "main_commit": lambda d: _flip(d["commitments"]["main_trace"][0]),
    "quotient_commit": lambda d: _flip(d["commitments"]["quotient_chunks"][5]),
    "pow_witness": lambda d: _flip(d["opening_proof"]["fri_proof"]["pow_witness"]),
    "final_poly": lambda d: _flip(d["opening_proof"]["fri_proof"]["final_poly"]["value"][0]),
    "fri_commit": lambda d: _flip(d["opening_proof"]["fri_proof"]["commit_phase_commits"][1][0]),
    "fri_sibling": lambda d: _flip(d["opening_proof"]["fri_proof"]["query_proofs"][5]["commit_phase_openings"][2]["sibling_value"]["value"][0]),
    "fri_path": lambda d: _flip(d["opening_proof"]["fri_proof"]["query_proofs"][0]["commit_phase_openings"][0]["opening_proof"][0][0]),
    "input_row": lambda d: _flip(d["opening_proof"]["query_openings"][3][0]["opened_values"][0][4]),
    "input_path": lambda d: _flip(d["opening_proof"]["query_openings"][7][2]["opening_proof"][1][3]),
    "log_degree": lambda d: d["chip_proofs"][3].__setitem__("log_degree", d["chip_proofs"][3]["log_degree"] + 1),
    "drop_query": lambda d: d["opening_proof"]["query_openings"].pop(),
    # one-row chips (height-2 LDEs, folded by no FRI round) and the heights of the chips with preprocessed columns
    "one_row_chip_trace": lambda d: _flip(d["chip_proofs"][13]["opened_values"]["trace_local"][2]["value"][0]),
    "one_row_chip_perm": lambda d: _flip(d["chip_proofs"][11]["opened_values"]["permutation_local"][0]["value"][0]),
    "one_row_chip_quotient": lambda d: _flip(d["chip_proofs"][9]["opened_values"]["quotient_chunks"][3]["value"][0]),
    "program_log_degree": lambda d: d["chip_proofs"][1].__setitem__("log_degree", 0),
    "range_log_degree": lambda d: d["chip_proofs"][12].__setitem__("log_degree", 0),
}


def test_one_row_chips_are_bound_by_the_opening_check(fib25):
    vb, cfg, t, proof = fib25
    for what in ("one_row_chip_trace", "one_row_chip_perm", "one_row_chip_quotient"):
        d = cbor2.loads(proof)
        TAMPERS[what](d)
        assert product_verdict(vb, cfg, cbor2.dumps(d), t.preprocessed) == -6, what      # VGPU_REJECT_FRI_FINAL, not a later check


@pytest.mark.parametrize("what", sorted(TAMPERS))
def test_rejects_tampering_like_the_oracle(fib25, oracle, what):
    vb, cfg, t, proof = fib25
    d = cbor2.loads(proof)
    TAMPERS[what](d)
    bad = cbor2.dumps(d)
    want = oracle.verify(bad, t.preprocessed)
    assert want != 0
    assert product_verdict(vb, cfg, bad, t.preprocessed) == expected(want)


def test_altered_witness_gets_the_oracle_verdict(fib25, oracle):
    """A witness edited after trace generation: whatever the prover and the oracle verifier make of it (the reference's
    CPU AIR does not tie the opcode flags to the opcode, so clearing one still proves), this verifier says the same."""
    vb, cfg, t, _ = fib25
    bad = vb.run_program(vb.fib_program(25), initial_fp=0x1000)
    row = next(i for i in range(bad.main[0].shape[0]) if bad.main[0][i, 22] == 1)
    bad.main[0][row, 22] = 0
    try:
        p = vb.prove_machine(cfg, bad)
    except vb.VgpuError:
        pytest.skip("prover refused the witness")
    assert product_verdict(vb, cfg, p, bad.preprocessed) == expected(oracle.verify(p, bad.preprocessed))


def test_malformed_bytes(fib25):
    vb, cfg, t, proof = fib25
    assert product_verdict(vb, cfg, proof[:-3], t.preprocessed) == -1
    assert product_verdict(vb, cfg, proof + b"\x00", t.preprocessed) == -1
    assert product_verdict(vb, cfg, b"", t.preprocessed) == -1
    d = cbor2.loads(proof)
    d["commitments"]["main_trace"][0]["value"] = 2013265921   # = p: not a field element
    assert product_verdict(vb, cfg, cbor2.dumps(d), t.preprocessed) == -1


def test_wrong_preprocessed_trace_rejected(fib25):
    vb, cfg, t, proof = fib25
    prep = [m.copy() for m in t.preprocessed]
    prep[0][0, 0] = (int(prep[0][0, 0]) + 1) % 2013265921
    assert product_verdict(vb, cfg, proof, prep) != 0
