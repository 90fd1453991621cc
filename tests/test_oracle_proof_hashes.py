"""The oracle prover's output bytes are pinned (tests/golden/oracle_proof_hashes.json): refactors of oracle/ — e.g. performance work
on the CPU baseline — must reproduce them exactly, for every thread count."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


def test_oracle_proofs_match_recorded_digests(built):
    import make_oracle_proof_hashes

    want = json.load(open(os.path.join(HERE, "golden", "oracle_proof_hashes.json")))
    assert make_oracle_proof_hashes.compute() == want


def test_oracle_proof_independent_of_thread_count(built, oracle):
    import hashlib
    import valida_b200 as vb

    t = vb.run_program(vb.fib_program(582), initial_fp=0x1000)
    want = json.load(open(os.path.join(HERE, "golden", "oracle_proof_hashes.json")))["fib_582"]["sha256"]
    top = oracle.max_threads()
    try:
        for th in (1, 3, top):
            oracle.set_threads(th)
            assert hashlib.sha256(oracle.prove(t.main, t.preprocessed, debug_checks=False).cbor()).hexdigest() == want, th
    finally:
        oracle.set_threads(top)
