"""Records SHA-256 digests of every trace matrix the host witness generator (valida_b200/csrc/host/tracegen.cc) produces
for a set of programs, as a regression fixture for changes to that generator (parallel sort, buffer handling):
    python tests/golden/make_trace_hashes.py        -> tests/golden/trace_hashes.json
The digests were first written by the single-threaded generator whose traces the oracle's constraint checker accepts
(tests/test_tracegen_and_oracle_prove.py); a later implementation must reproduce them bit for bit."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cases():
    import valida_b200 as vb
    import programs

    golden = json.load(open(os.path.join(HERE, "programs.json")))
    out = [("fib_%d" % n, vb.fib_program(n)) for n in (0, 1, 25, 582, 9360, 37446)]
    out += [(k, np.array(v["program"], dtype=np.int32)) for k, v in sorted(golden.items())]
    out += [("mixed_300", programs.mixed_program(300)), ("config5_2000", programs.config5_program(2000))]
    return out


def digest(t):
    d = {"clock": int(t.clock), "mem_ops": int(t.mem_ops), "add_ops": int(t.add_ops)}
    for i, m in enumerate(t.main):
        d["main_%d" % i] = [list(m.shape), hashlib.sha256(np.ascontiguousarray(m).tobytes()).hexdigest()]
    for i, m in enumerate(t.preprocessed):
        d["prep_%d" % i] = [list(m.shape), hashlib.sha256(np.ascontiguousarray(m).tobytes()).hexdigest()]
    return d


def compute():
    import valida_b200 as vb

    return {name: digest(vb.run_program(prog, initial_fp=0x1000)) for name, prog in cases()}


if __name__ == "__main__":
    json.dump(compute(), open(os.path.join(HERE, "trace_hashes.json"), "w"), indent=1, sort_keys=True)
    print("wrote", os.path.join(HERE, "trace_hashes.json"))
