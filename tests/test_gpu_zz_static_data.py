"""GPU parity for the reference's prove_static_data (basic/tests/test_static_data.rs:30-113): the static-data chip's trace has
real rows, the memory trace opens with the static rows, and the proof bytes are the oracle's.  (Runs last in the GPU suite.)"""
import pytest

import programs

pytestmark = pytest.mark.gpu


def test_prove_static_data_bytes_equal_oracle_and_verify(ctx, oracle):
    import valida_b200 as vb

    prog, cells = programs.static_data_program()
    t = vb.run_program(prog, initial_fp=0x1000, static_data=cells)
    assert t.clock == 4 and t.main[13].shape == (2, 6) and t.main[2].shape == (8, 14)
    cfg = vb.StarkConfig(ctx, oracle.rc480)
    proof = vb.prove_machine(cfg, t)
    ref = oracle.prove(t.main, t.preprocessed, debug_checks=False).cbor()
    assert proof == ref
    assert oracle.verify(proof, t.preprocessed) == 0        # machine.verify(&config, &proof).expect(..) of the reference test
    vb.verify_machine(cfg, proof, t.preprocessed)           # the library's own verifier


def test_static_data_proof_without_the_static_rows_is_rejected(ctx, oracle):
    """A prover that leaves the static-data chip's sends out cannot balance the memory bus: both verifiers reject."""
    import numpy as np
    import valida_b200 as vb

    prog, cells = programs.static_data_program()
    t = vb.run_program(prog, initial_fp=0x1000, static_data=cells)

    class Tampered:
        main = [m.copy() for m in t.main]
        preprocessed = t.preprocessed

    Tampered.main[13] = np.zeros_like(Tampered.main[13])
    cfg = vb.StarkConfig(ctx, oracle.rc480)
    proof = vb.prove_machine(cfg, Tampered)
    assert oracle.verify(proof, t.preprocessed) != 0
    with pytest.raises(vb.VerificationError):
        vb.verify_machine(cfg, proof, t.preprocessed)
