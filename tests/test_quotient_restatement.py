"""A second restatement, in plain Python integers, of the reference's IN-REPO quotient code for chips whose AIR is short enough
to transcribe here: quotient / quotient_values (machine/src/quotient.rs:18-238), ProverConstraintFolder::assert_zero
(machine/src/folding_builder.rs:62-66, 97-102: acc = acc * alpha + c), eval_permutation_constraints (machine/src/chip.rs:210-289)
and the AIRs of Add32Chip (alu_u32/src/add/stark.rs:21-55) and StaticDataChip (static_data/src/stark.rs:25-37); memory, range and
program have empty AIRs.  The oracle's prover and verifier share one constraint template, so "the verifier accepts" cannot see a
sign or order slip in it — this independent reading can.  The pieces that live in Plonky3 (selectors Z_H(x)/(x - g^i), the
even/odd chunk split) are written from their mathematical definitions [P3-UNVERIFIED, as in the oracle].
Random (constraint-violating) traces, 8 rows: every constraint contributes.  CPU only."""
import numpy as np
import pytest

from test_perm_trace_restatement import CHIPS, GENERAL, MEM, P, RANGE, RECEIVE, SEND, apply, e_add, e_from, e_inv, e_mul, e_sub  # noqa: F401

GEN = 31                       # coset shift = BabyBear multiplicative generator
TWO_ADIC_GEN_27 = 0x1a427a41   # two_adic_generator(27)


def two_adic_generator(bits):
    g = TWO_ADIC_GEN_27
    for _ in range(27 - bits):
        g = g * g % P
    return g


def e_scale(a, k): return [x * k % P for x in a]


# ---- AIRs: each returns the list of base-field constraint values in eval() order ---------------------------------------------
def air_add(loc, nxt, sel):
    """Add32Chip::eval (alu_u32/src/add/stark.rs:21-55): input_1 0-3, input_2 4-7, carry 8-10, output 11-14."""
    i1, i2, carry, out = loc[0:4], loc[4:8], loc[8:11], loc[11:15]
    base = 1 << 8
    ov = [(i1[3] + i2[3] - out[3]) % P, (i1[2] + i2[2] - out[2] + carry[0]) % P,
          (i1[1] + i2[1] - out[1] + carry[1]) % P, (i1[0] + i2[0] - out[0] + carry[2]) % P]
    cons = [o * (o - base) % P for o in ov]
    cons += [(ov[k] * (carry[k] - 1) + (ov[k] - base) * carry[k]) % P for k in range(3)]
    cons += [c * (c - 1) % P for c in carry]        # assert_bool
    return cons


def air_static_data(loc, nxt, sel):
    """StaticDataChip::eval_main (static_data/src/stark.rs:25-37): when_transition().when(local.is_real * next.is_real)
    .assert_eq(next.addr, local.addr + 4); addr 0, is_real 5."""
    cond = loc[5] * nxt[5] % P
    return [sel["transition"] * cond % P * ((nxt[0] - (loc[0] + 4)) % P) % P]


AIRS = {1: None, 2: None, 12: None, 3: air_add, 13: air_static_data}


def quotient_py(chip, log_degree, main_lde, perm_lde, cumsum, ch15, alpha):
    """main_lde / perm_lde: NATURAL-order evaluations over the coset GEN * <g_ext>, 2n rows."""
    n, qs = 1 << log_degree, 2 << log_degree
    g_sub, g_ext = two_adic_generator(log_degree), two_adic_generator(log_degree + 1)
    subgroup_last = pow(g_sub, P - 2, P)
    r1, r2 = [int(x) for x in ch15[5:10]], [int(x) for x in ch15[10:15]]
    alphas_global, acc = [], e_from(1)
    for _ in range(4):
        acc = e_mul(acc, r1)
        alphas_global.append(acc)
    inter = CHIPS[chip]
    k = len(inter)
    q = []
    x = GEN
    for i in range(qs):
        loc, nxt = [int(v) for v in main_lde[i]], [int(v) for v in main_lde[(i + 2) % qs]]      # next_step = 1 << log_quotient_degree
        pl = [[int(v) for v in perm_lde[i][5 * m:5 * m + 5]] for m in range(k + 1)]
        pn = [[int(v) for v in perm_lde[(i + 2) % qs][5 * m:5 * m + 5]] for m in range(k + 1)]
        zh = (pow(x, n, P) - 1) % P
        sel = {"first": zh * pow((x - 1) % P, P - 2, P) % P, "last": zh * pow((x - subgroup_last) % P, P - 2, P) % P,
               "transition": (x - subgroup_last) % P}
        folded = e_from(0)

        def assert_zero_ext(c):
            nonlocal folded
            folded = e_add(e_mul(folded, alpha), c)

        if AIRS[chip]:
            for c in AIRS[chip](loc, nxt, sel):
                assert_zero_ext(e_from(c))
        # eval_permutation_constraints (machine/src/chip.rs:210-289)
        phi_local, phi_next = pl[k], pn[k]
        lhs, rhs, phi0 = e_sub(phi_next, phi_local), e_from(0), e_from(0)
        for m, (sign, bus, fields, count) in enumerate(inter):
            rlc, beta = e_from(0), e_from(1)
            for f in fields:
                rlc = e_add(rlc, e_scale(beta, apply(f, loc)))
                beta = e_mul(beta, r2)
            rlc = e_add(rlc, alphas_global[bus])
            assert_zero_ext(e_sub(e_mul(rlc, pl[m]), e_from(1)))                 # assert_one_ext
            t_loc, t_nxt = e_scale(pl[m], apply(count, loc)), e_scale(pn[m], apply(count, nxt))
            if sign == SEND:
                phi0, rhs = e_add(phi0, t_loc), e_add(rhs, t_nxt)
            else:
                phi0, rhs = e_sub(phi0, t_loc), e_sub(rhs, t_nxt)
        assert_zero_ext(e_scale(e_sub(lhs, rhs), sel["transition"]))              # when_transition().assert_eq_ext(lhs, rhs)
        assert_zero_ext(e_scale(e_sub(phi_local, phi0), sel["first"]))            # when_first_row()
        assert_zero_ext(e_scale(e_sub(phi_local, cumsum), sel["last"]))           # when_last_row()
        q.append(e_scale(folded, pow(zh, P - 2, P)))
        x = x * g_ext % P
    # decompose_and_flatten, log_chunks = 1: q(x) = q_even(x^2) + x q_odd(x^2); -x_i = x_(i+n)
    out = np.zeros((n, 10), dtype=np.uint32)
    x, half = GEN, pow(2, P - 2, P)
    for i in range(n):
        out[i, :5] = e_scale(e_add(q[i], q[i + n]), half)
        out[i, 5:] = e_scale(e_sub(q[i], q[i + n]), half * pow(x, P - 2, P) % P)
        x = x * g_ext % P
    return out


@pytest.mark.parametrize("chip", sorted(AIRS))
def test_python_quotient_matches_oracle_on_random_traces(built, oracle, chip):
    rng = np.random.default_rng(300 + chip)
    log_degree = 3
    n = 1 << log_degree
    main = rng.integers(0, P, (n, oracle.chip_width(chip)), dtype=np.uint32)
    perm = rng.integers(0, P, (n, oracle.chip_perm_width(chip)), dtype=np.uint32)
    pw = oracle.chip_prep_width(chip)
    prep = rng.integers(0, P, (n, pw), dtype=np.uint32) if pw else None
    ch, alpha, cs = (rng.integers(0, P, k, dtype=np.uint32) for k in (15, 5, 5))
    nat = lambda m: oracle.coset_lde(m, 1, GEN, False)
    rev = lambda m: oracle.coset_lde(m, 1, GEN, True)
    want = quotient_py(chip, log_degree, nat(main), nat(perm), [int(v) for v in cs], ch, [int(v) for v in alpha])
    got = oracle.quotient(chip, log_degree, rev(prep) if pw else None, rev(main), rev(perm), cs, ch, alpha)
    assert np.array_equal(got, want)
