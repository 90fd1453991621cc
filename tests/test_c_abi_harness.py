"""tests/c/c_abi_smoke.c: a plain C caller of include/valida_b200.h (no Python, no torch in the process) — the checkable stand-in
for the Rust FFI crate of INTEGRATION.md.  CPU: it compiles against the header, links the library and fails LOUDLY without a GPU
(exit code 3, "no CPU fallback").  GPU: it proves fib(25), its own vgpu_verify accepts, and the bytes equal the Python path's."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "c_abi_smoke")
    lib_dir = os.path.join(ROOT, "valida_b200")
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "c_abi_smoke.c"),
                    "-o", exe, "-L", lib_dir, "-lvalida_b200", "-Wl,-rpath," + lib_dir], check=True)
    return exe


def test_c_caller_builds_and_refuses_to_run_without_a_gpu(built, tmp_path):
    import torch

    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run itself is covered by the gpu-marked test")
    r = subprocess.run([exe, "25"], capture_output=True, text=True)
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)
    assert "no CPU fallback" in r.stderr
    assert "192 cycles, 401 memory operations, 105 additions" in r.stdout      # basic/tests/test_prover.rs:479-482


@pytest.mark.gpu
def test_c_caller_proves_and_matches_the_python_path(built, ctx, oracle, tmp_path):
    import valida_b200 as vb

    exe = _build(tmp_path)
    out = str(tmp_path / "proof.cbor")
    r = subprocess.run([exe, "25", out], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "verdict 0" in r.stdout
    t = vb.run_program(vb.fib_program(25), initial_fp=0x1000)
    proof = open(out, "rb").read()
    assert proof == vb.prove_machine(vb.StarkConfig(ctx, oracle.rc480), t)
    assert oracle.verify(proof, t.preprocessed) == 0
