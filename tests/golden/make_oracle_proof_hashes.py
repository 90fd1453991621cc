"""SHA-256 of the oracle prover's CBOR output for a set of programs (fixed Poseidon constants: oracle_binding's rc480):
a regression pin for refactors of oracle/ (performance work on the CPU baseline must not change a byte).
    python tests/golden/make_oracle_proof_hashes.py   -> tests/golden/oracle_proof_hashes.json"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cases():
    import numpy as np
    import valida_b200 as vb
    import programs

    golden = json.load(open(os.path.join(HERE, "programs.json")))
    out = [("fib_%d" % n, vb.run_program(vb.fib_program(n), initial_fp=0x1000)) for n in (0, 3, 25, 582)]
    out += [(k, vb.run_program(np.array(v["program"], dtype=np.int32), initial_fp=0x1000)) for k, v in sorted(golden.items())]
    prog, cells = programs.static_data_program()
    out.append(("static_data", vb.run_program(prog, initial_fp=0x1000, static_data=cells)))
    out.append(("config5_40", vb.run_program(programs.config5_program(40), initial_fp=0x1000)))
    return out


def compute():
    import oracle_binding

    orc = oracle_binding.Oracle()
    res = {}
    for name, t in cases():
        proof = orc.prove(t.main, t.preprocessed, debug_checks=False).cbor()
        res[name] = {"bytes": len(proof), "sha256": hashlib.sha256(proof).hexdigest()}
    return res


if __name__ == "__main__":
    json.dump(compute(), open(os.path.join(HERE, "oracle_proof_hashes.json"), "w"), indent=1, sort_keys=True)
    print("wrote oracle_proof_hashes.json")
