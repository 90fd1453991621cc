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


# ---- VirtualPairCol: ("main", c) | ("const", k) | ("sum_main", [c...]) ------------------------------------------------
def apply(col, main_row):
    kind, v = col
    if kind == "main":
        return int(main_row[v])
    if kind == "const":
        return v % P
    return sum(int(main_row[c]) for c in v) % P


GENERAL, PROGRAM, MEM, RANGE = 0, 1, 2, 3      # basic/src/lib.rs:1190-1212
SEND, RECEIVE = +1, -1
word = lambda c0: [("main", c0 + i) for i in range(4)]

# all_interactions order: local sends, local receives, global sends, global receives (machine/src/chip.rs:40-63)
CHIPS = {
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
