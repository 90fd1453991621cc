"""A second, independent restatement of the reference's IN-REPO LogUp code — generate_permutation_trace
(machine/src/chip.rs:121-208), generate_rlc_elements (291-331), reduce_row (335-352),
batch_multiplicative_inverse_allowing_zero (util/src/lib.rs:21-43) — in plain Python integers, written from the Rust text
and from each chip's global_sends / global_receives, and compared with the C++ oracle on real witness traces.
The oracle (and through it the CUDA path) is thereby pinned against two separately written readings of the same source.
CPU only; small traces (pure-Python loops)."""
import numpy as np
import pytest

import programs

P = 2013265921


# ---- F_p[X]/(X^5 - 2) on Python ints -------------------------------------------------------------------------------
def e_add(a, b): return [(x + y) % P for x, y in zip(a, b)]
def e_sub(a, b): return [(x - y) % P for x, y in zip(a, b)]
def e_from(x): return [x % P, 0, 0, 0, 0]


def e_mul(a, b):
    t = [0] * 9
    for i in range(5):
        for j in range(5):
            t[i + j] += a[i] * b[j]
    return [(t[k] + 2 * (t[k + 5] if k + 5 < 9 else 0)) % P for k in range(5)]


def e_inv(a):
    """Solve (multiplication-by-a matrix) x = 1 by Gaussian elimination mod p — no shared code with the oracle's Frobenius inverse."""
    cols = []
    basis = [[1 if i == j else 0 for i in range(5)] for j in range(5)]
    for j in range(5):
        cols.append(e_mul(a, basis[j]))
    m = [[cols[j][i] for j in range(5)] + [1 if i == 0 else 0] for i in range(5)]
    for c in range(5):
        piv = next(r for r in range(c, 5) if m[r][c] % P)
        m[c], m[piv] = m[piv], m[c]
        inv = pow(m[c][c], P - 2, P)
        m[c] = [x * inv % P for x in m[c]]
        for r in range(5):
            if r != c and m[r][c]:
                f = m[r][c]
                m[r] = [(x - f * y) % P for x, y in zip(m[r], m[c])]
    return [m[i][5] for i in range(5)]


# ---- VirtualPairCol: ("main", c) | ("const", k) | ("sum_main", [c...]) | ("weighted", [(c, weight)...]) -------------
def apply(col, main_row):
    kind, v = col
    if kind == "main":
        return int(main_row[v])
    if kind == "const":
        return v % P
    if kind == "weighted":                  # VirtualPairCol::new_main(vec![(col, weight)...], zero)
        return sum(int(main_row[c]) * w for c, w in v) % P
    return sum(int(main_row[c]) for c in v) % P


GENERAL, PROGRAM, MEM, RANGE = 0, 1, 2, 3      # basic/src/lib.rs:1190-1212
SEND, RECEIVE = +1, -1
word = lambda c0: [("main", c0 + i) for i in range(4)]

zero3_then = lambda c: [("const", 0)] * 3 + [("main", c)]


def cpu_interactions():
    """CpuChip::global_sends (cpu/src/lib.rs:99-159).  CpuCols (cpu/src/columns.rs:8-77): clk 0, pc 1, fp 2, opcode 3,
    operands 4-8, 17 opcode flags 9-25 (is_bus_op first), diff 26, diff_inv 27, not_equal 28, three memory channels of
    (used, is_read, addr, value[4]) at 29 / 36 / 43, chip_channel.clk_or_zero 50."""
    out = []
    for ch in (29, 36, 43):
        out.append((SEND, MEM, [("main", ch + 1), ("main", 0), ("main", ch + 2), ("const", 0)] + word(ch + 3), ("main", ch)))
    out.append((SEND, GENERAL, [("main", 3)] + word(32) + word(39) + word(46) + [("main", 50)], ("main", 9)))
    return out


def alu(opcode, in1, in2, out, count):
    return (RECEIVE, GENERAL, [opcode] + word(in1) + word(in2) + out, count)


# all_interactions order: local sends, local receives, global sends, global receives (machine/src/chip.rs:40-63)
CHIPS = {
    0: cpu_interactions(),
    1: [],   # ProgramChip: no interactions (program/src/lib.rs:50-68; the program-bus send of the CPU is commented out)
    # Sub32Chip (alu_u32/src/sub/mod.rs:53-88): the Add32 layout with `borrow` for `carry`, opcode SUB32 = 101
    4: [(SEND, RANGE, [("main", 11 + i)], ("main", 15)) for i in range(4)]
       + [(RECEIVE, GENERAL, [("const", 101)] + word(0) + word(4) + word(11), ("main", 15))],
    # Mul32Chip (mul/mod.rs:66-101; columns: in1 0-3, in2 4-7, out 8-11, r 12, s 13, is_mul 14, is_mulhs 15, is_mulhu 16, counter 17)
    5: [alu(("weighted", [(14, 102), (15, 114), (16, 112)]), 0, 4, word(8), ("sum_main", [14, 15, 16]))],
    # Div32Chip (div/mod.rs:55-80; in1 0-3, in2 4-7, out 8-11, is_div 12, is_sdiv 13)
    6: [alu(("weighted", [(12, 103), (13, 110)]), 0, 4, word(8), ("sum_main", [12, 13]))],
    # Shift32Chip (shift/mod.rs:58-116; in1 0-3, in2 4-7, out 8-11, bits_2 12-19, temp_1 20, power_of_two 21-24, is_shl 25, is_shr 26, is_sra 27):
    # sends a MUL32 / DIV32 / SDIV32 with the power of two, receives the SHL32 / SHR32 / SRA32
    7: [(SEND, GENERAL, [("weighted", [(25, 102), (26, 103), (27, 110)])] + word(0) + word(21) + word(8), ("sum_main", [25, 26, 27])),
        alu(("weighted", [(25, 105), (26, 106), (27, 113)]), 0, 4, word(8), ("sum_main", [25, 26, 27]))],
    # Lt32Chip (lt/mod.rs:58-85; in1 0-3, in2 4-7, byte_flag 8-11, bits 12-20, output 21, multiplicity 22, is_lt 23, is_lte 24, is_slt 25, is_sle 26)
    8: [alu(("weighted", [(23, 104), (24, 115), (25, 117), (26, 118)]), 0, 4, zero3_then(21), ("main", 22))],
    # Com32Chip (com/mod.rs:56-83; in1 0-3, in2 4-7, diff 8, diff_inv 9, not_equal 10, output 11, is_ne 12, is_eq 13)
    9: [alu(("weighted", [(12, 111), (13, 116)]), 0, 4, zero3_then(11), ("sum_main", [12, 13]))],
    # Bitwise32Chip (bitwise/mod.rs:56-82; in1 0-3, in2 4-7, bits_1 8-39, bits_2 40-71, output 72-75, is_and 76, is_or 77, is_xor 78)
    10: [alu(("weighted", [(76, 107), (77, 108), (78, 109)]), 0, 4, word(72), ("sum_main", [76, 77, 78]))],
    # OutputChip (output/src/lib.rs:117-136; clk 0, value 1, is_real 2, diff 3, counter 4, counter_mult 5, opcode 6):
    # (opcode, twelve value bytes of which only the fourth is the output byte, clk)
    11: [(RECEIVE, GENERAL, [("main", 6)] + [("main", 1) if i == 3 else ("const", 0) for i in range(12)] + [("main", 0)], ("main", 2))],
    # Add32Chip (alu_u32/src/add/mod.rs:53-88; columns.rs: input_1 0-3, input_2 4-7, carry 8-10, output 11-14, is_real 15)
    3: [(SEND, RANGE, [("main", 11 + i)], ("main", 15)) for i in range(4)]
       + [(RECEIVE, GENERAL, [("const", 100)] + word(0) + word(4) + word(11), ("main", 15))],
    # MemoryChip (memory/src/lib.rs:216-233; columns.rs: addr 0, value 1-4, clk 5, is_static_initial 6, is_read 7, is_write 8)
    2: [(RECEIVE, MEM, [("main", 7), ("main", 5), ("main", 0), ("main", 6)] + word(1), ("sum_main", [7, 8]))],
    # RangeCheckerChip (range/src/lib.rs:45-56; columns.rs: mult 0, counter 1)
    12: [(RECEIVE, RANGE, [("main", 1)], ("main", 0))],
    # StaticDataChip (static_data/src/lib.rs:81-96; columns.rs: addr 0, value 1-4, is_real 5)
    13: [(SEND, MEM, [("const", 0), ("const", 0), ("main", 0), ("const", 1)] + word(1), ("main", 5))],
}


def perm_trace_py(main, interactions, ch15):
    r1, r2 = [int(x) for x in ch15[5:10]], [int(x) for x in ch15[10:15]]
    alphas_global, acc = [], e_from(1)
    for _ in range(4):                                  # powers().skip(1): alphas_global[i] = r1^(i+1)
        acc = e_mul(acc, r1)
        alphas_global.append(acc)
    h, k = main.shape[0], len(interactions)
    denoms = []
    for n in range(h):
        for (_, bus, fields, _) in interactions:
            rlc, beta = e_from(0), e_from(1)
            for f in fields:                            # reduce_row: rlc += beta^j * field_j, then += alpha
                rlc = e_add(rlc, e_mul(beta, e_from(apply(f, main[n]))))
                beta = e_mul(beta, r2)
            denoms.append(e_add(rlc, alphas_global[bus]))
    q = [d if not any(d) else e_inv(d) for d in denoms]  # zero stays zero (util/src/lib.rs:21-43)
    out = np.zeros((h, 5 * (k + 1)), dtype=np.uint32)
    phi = e_from(0)
    for n in range(h):
        for m, (sign, _, _, count) in enumerate(interactions):
            qm = q[n * k + m]
            out[n, 5 * m:5 * m + 5] = qm
            term = e_mul(qm, e_from(apply(count, main[n])))
            phi = e_add(phi, term) if sign == SEND else e_sub(phi, term)
        out[n, 5 * k:5 * k + 5] = phi
    return out, phi


def test_ext5_inverse_by_elimination():
    rng = np.random.default_rng(1)
    for _ in range(20):
        a = [int(x) for x in rng.integers(0, P, 5)]
        assert e_mul(a, e_inv(a)) == [1, 0, 0, 0, 0]


@pytest.mark.parametrize("chip", [3, 2, 12])
def test_python_restatement_matches_oracle_on_fib_traces(built, oracle, chip):
    import valida_b200 as vb

    t = vb.run_program(vb.fib_program(3), initial_fp=0x1000)        # 38 cycles: cpu 64 rows, mem 128, add 32, range 256
    ch = np.random.default_rng(100 + chip).integers(0, P, 15, dtype=np.uint32)
    main = t.main[chip]
    prep = t.preprocessed[1] if chip == 12 else None
    want, cs = perm_trace_py(main, CHIPS[chip], ch)
    got, got_cs = oracle.perm_trace(chip, main, prep, ch)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    assert [int(x) for x in got_cs] == cs


@pytest.mark.parametrize("chip", sorted(CHIPS))
def test_interaction_tables_of_every_chip_on_random_traces(built, oracle, chip):
    """Random (non-witness) rows make every field, weight and count matter: a wrong column, opcode weight, bus or direction in
    the oracle's chip table (the table the CUDA path is compared with) changes the permutation trace."""
    rng = np.random.default_rng(1000 + chip)
    h, w = 8, oracle.chip_width(chip)
    main = rng.integers(0, P, (h, w), dtype=np.uint32)
    pw = oracle.chip_prep_width(chip)
    prep = rng.integers(0, P, (h, pw), dtype=np.uint32) if pw else None
    ch = rng.integers(0, P, 15, dtype=np.uint32)
    want, cs = perm_trace_py(main, CHIPS[chip], ch)
    got, got_cs = oracle.perm_trace(chip, main, prep, ch)
    assert oracle.chip_perm_width(chip) == 5 * (len(CHIPS[chip]) + 1)
    assert np.array_equal(got, want)
    assert [int(x) for x in got_cs] == cs


def test_python_restatement_matches_oracle_on_static_data(built, oracle):
    import valida_b200 as vb

    prog, cells = programs.static_data_program()
    t = vb.run_program(prog, initial_fp=0x1000, static_data=cells)
    ch = np.random.default_rng(7).integers(0, P, 15, dtype=np.uint32)
    sums = []
    for chip in (13, 2):
        want, cs = perm_trace_py(t.main[chip], CHIPS[chip], ch)
        got, got_cs = oracle.perm_trace(chip, t.main[chip], None, ch)
        assert np.array_equal(got, want)
        assert [int(x) for x in got_cs] == cs
        sums.append(cs)
    # the static rows of the memory trace receive exactly what the static-data chip sends: after removing the CPU's memory
    # traffic (its sends are not in these two sums) the static contributions cancel pairwise — check them directly
    r1, r2 = [int(x) for x in ch[5:10]], [int(x) for x in ch[10:15]]
    _, only_static_rows = perm_trace_py(t.main[2][:2], CHIPS[2], ch)
    assert e_add(sums[0], only_static_rows) == [0, 0, 0, 0, 0]
