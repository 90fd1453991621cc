#!/usr/bin/env python
"""Writes the inputs of the parity experiment (rust/parity/README.md): fib25.words and expected.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "."
    os.makedirs(out, exist_ok=True)
    import valida_b200 as vb

    prog = vb.fib_program(25)
    with open(os.path.join(out, "fib25.words"), "w") as f:
        for row in prog:
            f.write(" ".join(str(int(x)) for x in row) + "\n")
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_proof_hashes.json")))["fib_25"]
    json.dump({"program": "fib n=25, fp = 0x1000", "proof_bytes": golden["bytes"], "sha256": golden["sha256"],
               "poseidon_round_constants": "SplitMix64(0x76616c696461), z >> 33 kept when < p, 480 words; CosetMds::default()"},
              open(os.path.join(out, "expected.json"), "w"), indent=1)
    print("wrote", os.path.join(out, "fib25.words"), "and expected.json:", golden["bytes"], "bytes, sha256", golden["sha256"])


if __name__ == "__main__":
    main()
