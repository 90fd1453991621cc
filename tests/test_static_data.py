"""The reference's fifth proving test, prove_static_data (basic/tests/test_static_data.rs:30-113): static memory preloaded
through the static-data chip, whose rows open the memory trace and balance the memory bus.  CPU only: host witness
generation + the oracle prover / verifier with the reference's debug-build invariants."""
import numpy as np
import pytest

import programs

P = 2013265921


@pytest.fixture(scope="module")
def static_run(built):
    import valida_b200 as vb

    prog, cells = programs.static_data_program()
    return vb.run_program(prog, initial_fp=0x1000, static_data=cells)


def test_static_program_terminates_and_trace_shapes(static_run):
    t = static_run
    assert t.clock == 4                       # imm32, load32, bnei (not taken: the loaded cell equals 0x25), stop
    assert t.mem_ops == 5                     # imm32 write; load32 read, read, write; bnei read
    assert t.mem_cell(0x1000 - 4) == 0x25     # the static value reached the frame
    assert t.mem_cell(0x14) == 0x32
    sd = t.main[13]
    assert sd.shape == (2, 6)                 # StaticDataCols: addr, value[4], is_real (static_data/src/columns.rs:8-17)
    assert sd.tolist() == [[0x10, 0, 0, 0, 0x25, 1], [0x14, 0, 0, 0, 0x32, 1]]
    mem = t.main[2]
    assert mem.shape == (8, 14)               # 2 static rows + 5 operations, padded
    # static rows first: is_static_initial = 1, clk = 0, is_write = 1, counter = n (memory/src/lib.rs:265-283)
    assert mem[0].tolist() == [0x10, 0, 0, 0, 0x25, 0, 1, 0, 1, 0, 0, 0, 0, 0]
    assert mem[1].tolist() == [0x14, 0, 0, 0, 0x32, 0, 1, 0, 1, 0, 0, 0, 1, 0]
    assert mem[2:7, 6].tolist() == [0] * 5 and mem[2:7, 12].tolist() == [2, 3, 4, 5, 6]
    assert (mem[7] == 0).all()


def test_without_static_data_the_load_faults(built):
    import valida_b200 as vb

    prog, _ = programs.static_data_program()
    with pytest.raises(vb.VgpuError, match="read before write"):
        vb.run_program(prog, initial_fp=0x1000)


def test_repeated_static_address_keeps_last_value_and_order_is_by_address(built):
    import valida_b200 as vb

    prog, _ = programs.static_data_program()
    t = vb.run_program(prog, initial_fp=0x1000, static_data={0x20: 7, 0x10: 0x25, 0x18: 0x01020304})
    assert t.main[13].tolist() == [[0x10, 0, 0, 0, 0x25, 1], [0x18, 1, 2, 3, 4, 1], [0x20, 0, 0, 0, 7, 1], [0, 0, 0, 0, 0, 0]]


def test_oracle_proves_and_verifies_static_data(static_run, oracle):
    t = static_run
    pr = oracle.prove(t.main, t.preprocessed, debug_checks=True)
    # the reference's debug-build invariants: every constraint vanishes on every row, the bus sums cancel — the memory bus
    # balances only if the static rows of the memory trace match the static-data chip's sends
    assert pr.constraint_failures() == [-1] * 14
    assert pr.cumulative_sum_zero()
    proof = pr.cbor()
    assert oracle.verify(proof, t.preprocessed) == 0


def test_static_rows_are_needed_for_the_bus_to_balance(static_run, oracle):
    t = static_run
    main = [m.copy() for m in t.main]
    main[13] = np.zeros_like(main[13])        # drop the static-data chip's sends
    pr = oracle.prove(main, t.preprocessed, debug_checks=True)
    assert not pr.cumulative_sum_zero()
