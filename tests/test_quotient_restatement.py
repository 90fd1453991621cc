"""A second restatement, in plain Python integers, of the reference's IN-REPO quotient code and of the AIRs of
all chips: quotient / quotient_values (machine/src/quotient.rs:18-238), ProverConstraintFolder::assert_zero
(machine/src/folding_builder.rs:62-66, 97-102: acc = acc * alpha + c), eval_permutation_constraints (machine/src/chip.rs:210-289)
and the AIRs of ALL fourteen chips (each air_* function cites its eval(); memory, range, program and div have empty AIRs).
The oracle's prover and verifier share one constraint template, so "the verifier accepts" cannot see a sign or order slip
in it — this independent reading can.  The pieces that live in Plonky3 (selectors Z_H(x)/(x - g^i), the
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


class X:
    """An element of F_p[X]/(X^5 - 2) with int-like operators (ints embed as constants): lets one AIR text serve the prover
    folder (base-field cells: plain ints) and the verifier folder (cells opened at zeta: extension values)."""
    __slots__ = ("c",)

    def __init__(self, c): self.c = [int(v) % P for v in c]
    @staticmethod
    def of(v): return v if isinstance(v, X) else X([v, 0, 0, 0, 0])
    def __add__(self, o): return X(e_add(self.c, X.of(o).c))
    __radd__ = __add__
    def __sub__(self, o): return X(e_sub(self.c, X.of(o).c))
    def __rsub__(self, o): return X(e_sub(X.of(o).c, self.c))
    def __mul__(self, o): return X(e_mul(self.c, X.of(o).c))
    __rmul__ = __mul__
    def __mod__(self, m): return self                  # the int-flavoured AIR text reduces with "% P": a no-op here
    def __eq__(self, o): return self.c == X.of(o).c
    def inv(self): return X(e_inv(self.c))


# ---- AIRs: each returns the list of base-field constraint values in eval() order ---------------------------------------------
def air_add(loc, nxt, sel):
    """Add32Chip::eval (alu_u32/src/add/stark.rs:21-55): input_1 0-3, input_2 4-7, carry 8-10, output 11-14."""
    i1, i2, carry, out = loc[0:4], loc[4:8], loc[8:11], loc[11:15]
    base = 1 << 8
    ov = [i1[3] + i2[3] - out[3], i1[2] + i2[2] - out[2] + carry[0], i1[1] + i2[1] - out[1] + carry[1], i1[0] + i2[0] - out[0] + carry[2]]
    cons = [o * (o - base) % P for o in ov]
    cons += [(ov[k] * (carry[k] - 1) + (ov[k] - base) * carry[k]) % P for k in range(3)]
    cons += [c * (c - 1) % P for c in carry]        # assert_bool
    return cons


def air_static_data(loc, nxt, sel):
    """StaticDataChip::eval_main (static_data/src/stark.rs:25-37): when_transition().when(local.is_real * next.is_real)
    .assert_eq(next.addr, local.addr + 4); addr 0, is_real 5."""
    cond = loc[5] * nxt[5]
    return [sel["transition"] * cond * (nxt[0] - (loc[0] + 4)) % P]


def air_cpu(loc, nxt, sel):
    """CpuChip::eval (cpu/src/stark.rs:22-305), transcribed in eval() order: eval_pc, eval_fp, eval_equality,
    eval_memory_channels, clock, immediates, stop.  CpuCols (cpu/src/columns.rs): clk 0, pc 1, fp 2, opcode 3, operands a..e 4-8,
    flags 9-25 in declaration order, diff 26, diff_inv 27, not_equal 28, channels (used, is_read, addr, value[4]) at 29 / 36 / 43,
    clk_or_zero 50."""
    tr, first, last = sel["transition"], sel["first"], sel["last"]
    clk, pc, fp = loc[0], loc[1], loc[2]
    opa, opb, opc, opd, ope = loc[4:9]
    (is_bus_op, is_bus_op_with_mem, is_imm_op, is_left_imm_op, is_load, _lu8, _ls8, is_store, _su8, is_beq, is_bne, is_jal, is_jalv,
     is_imm32, is_advice, is_stop, is_loadfp) = loc[9:26]
    diff, diff_inv, not_equal = loc[26], loc[27], loc[28]
    r1_used, r1_is_read, r1_addr, rv1 = loc[29], loc[30], loc[31], loc[32:36]
    r2_used, r2_is_read, r2_addr, rv2 = loc[36], loc[37], loc[38], loc[39:43]
    w_used, w_is_read, w_addr, wv = loc[43], loc[44], loc[45], loc[46:50]
    clk_or_zero = loc[50]
    npc, nfp, nclk = nxt[1], nxt[2], nxt[0]
    base = [1 << 24, 1 << 16, 1 << 8, 1]
    reduce = lambda word: sum(b * x for b, x in zip(base, word))
    sqdiff = lambda u, v: sum((a - b) * (a - b) for a, b in zip(u, v))
    B = 24                                            # BYTES_PER_INSTR
    cons = []
    z = _z(cons)
    # eval_pc (stark.rs:206-251)
    should_inc = is_imm32 + is_loadfp + is_bus_op + is_advice
    inc_pc = pc + 1
    z(tr, should_inc, npc - inc_pc)
    equal = 1 - not_equal
    beq_next = equal * opa + B * not_equal * inc_pc
    bne_next = B * equal * inc_pc + not_equal * opa
    z(tr, is_beq, B * npc - beq_next)
    z(tr, is_bne, B * npc - bne_next)
    z(tr, is_jal, B * npc - opb)
    z(tr, is_jalv, B * npc - reduce(rv1))
    # eval_fp (253-274)
    z(tr, is_jal, nfp - (fp + opc))
    z(tr, is_jalv, nfp - (fp + reduce(rv2)))
    z(tr, 1 - is_jal - is_jalv, nfp - fp)
    # eval_equality (276-297)
    z(diff - sqdiff(rv1, rv2))
    z(not_equal, not_equal - 1)
    z(not_equal - diff * diff_inv)
    z(equal * diff)
    # eval_memory_channels (69-204)
    for flag in (is_load, is_store, is_jal, is_jalv, is_beq, is_bne, is_imm32, is_loadfp, is_imm_op, is_left_imm_op, is_bus_op):
        z(flag, flag - 1)
    addr_a, addr_b, addr_c = fp + opa, fp + opb, fp + opc
    z(r1_is_read - 1)
    z(r2_is_read - 1)
    z(w_is_read)
    z(is_jalv + is_beq + is_bne + is_bus_op * (1 - is_left_imm_op), r1_addr - addr_b)
    z(is_load + is_store, r1_addr - addr_c)
    z(is_load + is_store + is_jalv + is_beq + is_bne + (1 - is_left_imm_op) * is_bus_op, r1_used - 1)
    z(is_jal + is_left_imm_op + is_loadfp + is_imm32, r1_used)
    z(is_load, r2_addr - reduce(rv1))
    z(is_store, r2_addr - addr_b)
    z(is_jalv + (1 - is_imm_op) * is_bus_op, r2_addr - addr_c)
    z(is_load + is_store + is_jalv + (1 - is_imm_op) * (is_beq + is_bne + is_bus_op), r2_used - 1)
    z(is_jal + is_imm_op * (is_beq + is_bne + is_bus_op) + is_loadfp + is_imm32, r2_used)
    z(is_load + is_jal + is_jalv + is_imm32 + is_bus_op + is_loadfp, w_addr - addr_a)
    z(is_store, w_addr - reduce(rv2))
    z(is_store, sqdiff(rv1, wv))
    z(is_load, sqdiff(rv2, wv))
    z(tr, is_jal + is_jalv, B * (pc + 1) - reduce(wv))
    z(is_imm32, sqdiff(wv, [opb, opc, opd, ope]))
    z(is_loadfp, addr_b - reduce(wv))
    z(is_store + is_load + is_jal + is_jalv + is_imm32 + is_loadfp + is_bus_op, w_used - 1)
    z(is_beq + is_bne, w_used)
    # clock (33-43)
    z(first, clk)
    z(tr, clk + 1 - nclk)
    z(is_bus_op_with_mem, clk - clk_or_zero)
    z(1 - is_bus_op_with_mem, clk_or_zero)
    # immediates (48-56)
    s_imm = is_imm_op + is_left_imm_op
    z(s_imm, s_imm - 1)
    z(is_imm_op, opc - reduce(rv2))
    z(is_left_imm_op, opb - reduce(rv1))
    # stop (59-65)
    z(tr, is_stop, npc - pc)
    z(last, is_stop - 1)
    return cons


def _z(cons):
    def z(*factors):
        v = 1
        for f in factors:
            v = v * f
        cons.append(v % P)
    return z


def bits_value(bits):
    """sum_k bit_k * 2^k without shifting (cells may be extension values)"""
    return sum(b * (1 << k) for k, b in enumerate(bits))


def air_sub(loc, nxt, sel):
    """Sub32Chip::eval (alu_u32/src/sub/stark.rs:21-52): input_1 0-3, input_2 4-7, borrow 8-10, output 11-14."""
    i1, i2, br, out = loc[0:4], loc[4:8], loc[8:11], loc[11:15]
    cons = []
    z = _z(cons)
    z(out[3] - (256 * br[0] + i1[3] - i2[3]))
    z(out[2] - (256 * br[1] + i1[2] - i2[2] - br[0]))
    z(out[1] - (256 * br[2] + i1[1] - i2[1] - br[1]))
    z(out[0] - (i1[0] - i2[0] - br[2]))
    for b in br:
        z(b, b - 1)
    return cons


def air_mul(loc, nxt, sel):
    """Mul32Chip::eval (alu_u32/src/mul/stark.rs:23-82): input_1 0-3, input_2 4-7, output 8-11, r 12, s 13, counter 17."""
    i1, i2, out, r, s_, counter, ncounter = loc[0:4], loc[4:8], loc[8:12], loc[12], loc[13], loc[17], nxt[17]
    base_m = [1, 1 << 8, 1 << 16, 1 << 24]
    pi_m = lambda N: sum(base_m[i + j] * i1[3 - i] * i2[3 - j] for i in range(N) for j in range(N) if i + j < N)
    sigma_m = lambda N: sum(base_m[i] * x for i, x in enumerate(list(reversed(out))[:N]))
    cons = []
    z = _z(cons)
    z(pi_m(4) - sigma_m(4) - r * 2)
    z(pi_m(2) - sigma_m(2) - s_ * base_m[2])
    z(sel["first"], counter - 1)
    cd = ncounter - counter
    z(sel["transition"], cd * (cd - 1))
    z(sel["last"], counter - (1 << 10))
    return cons


def air_shift(loc, nxt, sel):
    """Shift32Chip::eval (alu_u32/src/shift/stark.rs:21-70): input_2 4-7, bits_2 12-19, temp_1 20, power_of_two 21-24, is_shl 25, is_shr 26, is_sra 27."""
    i2, bits, temp_1, pw2, shl, shr, sra = loc[4:8], loc[12:20], loc[20], loc[21:25], loc[25], loc[26], loc[27]
    cons = []
    z = _z(cons)
    z(i2[3] - bits_value(bits))
    for b in bits:
        z(b, b - 1)
    z(temp_1 - (bits[0] * 2) * (bits[1] * 4) * (bits[2] * 16))
    z(pw2[0] - temp_1 * (1 - bits[3]) * (1 - bits[4]))
    z(pw2[1] - temp_1 * bits[3] * (1 - bits[4]))
    z(pw2[2] - temp_1 * (1 - bits[3]) * bits[4])
    z(pw2[3] - temp_1 * bits[3] * bits[4])
    for f in (shl, shr, sra, shl + shr + sra):
        z(f, f - 1)
    return cons


def air_lt(loc, nxt, sel):
    """Lt32Chip::eval (alu_u32/src/lt/stark.rs:21-169): input_1 0-3, input_2 4-7, byte_flag 8-11, bits 12-20, output 21, multiplicity 22,
    is_lt 23, is_lte 24, is_slt 25, is_sle 26, diff_inv 27, top_bits_1 28-35, top_bits_2 36-43, different_signs 44.
    when_ne(a, b) filters with (a - b)."""
    i1, i2, bf, bits, out = loc[0:4], loc[4:8], loc[8:12], loc[12:21], loc[21]
    is_lt, is_lte, is_slt, is_sle, diff_inv = loc[23], loc[24], loc[25], loc[26], loc[27]
    tb1, tb2, dsig = loc[28:36], loc[36:44], loc[44]
    cons = []
    z = _z(cons)
    bit_comp = bits_value(bits)
    flag_sum = sum(bf)
    z(flag_sum, flag_sum - 1)
    z(bf[0] - 1, i1[0] - i2[0])
    z(bf[0] + bf[1] - 1, i1[1] - i2[1])
    z(bf[0] + bf[1] + bf[2] - 1, i1[2] - i2[2])
    z(flag_sum - 1, i1[3] - i2[3])
    z(flag_sum - 1, bit_comp)
    for i in range(4):
        z(bf[i], 256 + i1[i] - i2[i] - bit_comp)
        z(bf[i], (i1[i] - i2[i]) * diff_inv - 1)
        z(bf[i], bf[i] - 1)
    z(bits_value(tb1) - i1[0])
    z(bits_value(tb2) - i2[0])
    is_signed = is_slt + is_sle
    is_unsigned, same_sign, are_equal = 1 - is_signed, 1 - dsig, 1 - flag_sum
    z(is_unsigned, dsig)
    z(is_signed, tb1[7] - tb2[7], dsig - 1)
    z(dsig, bf[0] - 1)
    z(dsig, tb1[7] + tb2[7] - 1)
    for f in (is_lt, is_lte, is_slt, is_sle, is_lt + is_lte + is_slt + is_sle):
        z(f, f - 1)
    z(bits[8], is_unsigned + same_sign, out)
    z(bits[8], dsig, out - 1)
    z(bits[8] + are_equal - 1, is_unsigned + same_sign, out - 1)
    z(bits[8] + are_equal - 1, dsig, out)
    z(are_equal, is_lte + is_sle, out - 1)
    z(are_equal, is_lt + is_slt, out)
    for b in list(bits) + list(tb1) + list(tb2):
        z(b, b - 1)
    return cons


def air_com(loc, nxt, sel):
    """Com32Chip::eval (alu_u32/src/com/stark.rs:21-50): input_1 0-3, input_2 4-7, diff 8, diff_inv 9, not_equal 10, output 11, is_ne 12, is_eq 13."""
    i1, i2, diff, diff_inv, ne, out, is_ne, is_eq = loc[0:4], loc[4:8], loc[8], loc[9], loc[10], loc[11], loc[12], loc[13]
    cons = []
    z = _z(cons)
    z(diff - sum((a - b) * (a - b) for a, b in zip(i1, i2)))
    z(ne, ne - 1)
    z(ne - diff * diff_inv)
    z((1 - ne) * diff)
    for f in (is_ne, is_eq, is_ne + is_eq):
        z(f, f - 1)
    z(out - (is_ne * ne + is_eq * (1 - ne)))
    return cons


def air_bitwise(loc, nxt, sel):
    """Bitwise32Chip::eval (alu_u32/src/bitwise/stark.rs:22-75): input_1 0-3, input_2 4-7, bits_1[byte][bit] 8 + 8 byte + bit, bits_2 40 + ...,
    output 72-75, is_and 76, is_or 77, is_xor 78."""
    i1, i2, out, is_and, is_or, is_xor = loc[0:4], loc[4:8], loc[72:76], loc[76], loc[77], loc[78]
    cons = []
    z = _z(cons)
    for i in range(4):
        b1, b2 = loc[8 + 8 * i:16 + 8 * i], loc[40 + 8 * i:48 + 8 * i]
        byte_1, byte_2 = bits_value(b1), bits_value(b2)
        z(i1[i] - byte_1)
        z(i2[i] - byte_2)
        b_and = sum(x * y * (1 << k) for k, (x, y) in enumerate(zip(b1, b2)))
        z(is_and, b_and - out[i])
        z(is_or, byte_1 + byte_2 - b_and - out[i])
        z(is_xor, byte_1 + byte_2 - 2 * b_and - out[i])
        for b in list(b1) + list(b2):
            z(b, b - 1)
    for f in (is_and, is_or, is_xor, is_and + is_or + is_xor):
        z(f, f - 1)
    return cons


def air_output(loc, nxt, sel):
    """OutputChip::eval (output/src/stark.rs:21-39): clk 0, value 1, is_real 2, diff 3, counter 4, counter_mult 5, opcode 6; WRITE = 300."""
    cons = []
    z = _z(cons)
    z(sel["transition"], loc[3] - (nxt[0] - loc[0]))
    z(sel["transition"], nxt[4] - (loc[4] + 1))
    z(loc[2], loc[6] - 300)
    return cons


# memory, range, program, div: empty eval bodies (memory/src/stark.rs, range/src/stark.rs, program/src/stark.rs, alu_u32/src/div/stark.rs)
AIRS = {0: air_cpu, 1: None, 2: None, 3: air_add, 4: air_sub, 5: air_mul, 6: None, 7: air_shift, 8: air_lt, 9: air_com, 10: air_bitwise,
        11: air_output, 12: None, 13: air_static_data}


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


# ---- verify_constraints (machine/src/verify.rs:11-107) on the opened values of a real proof -------------------------------------
def verify_constraints_py(chip, log_degree, trace_local, trace_next, perm_local, perm_next, quotient_chunks, cumsum, zeta, alpha, ch15):
    """Every argument is a list of ext5 coefficient lists, as they sit in OpenedValues.  Returns (folded constraints, z_h * quotient)."""
    n = 1 << log_degree
    g_inv = pow(two_adic_generator(log_degree), P - 2, P)
    zeta = X(zeta)
    zeta_n = zeta
    for _ in range(log_degree):
        zeta_n = zeta_n * zeta_n
    z_h = zeta_n - 1
    sel = {"first": z_h * (zeta - 1).inv(), "last": z_h * (zeta - g_inv).inv(), "transition": zeta - g_inv}
    monomial = lambda k: X([1 if i == k else 0 for i in range(5)])
    unflatten = lambda v: [sum(X(v[5 * m + k]) * monomial(k) for k in range(5)) for m in range(len(v) // 5)]   # sum_k x_k * X^k
    loc, nxt = [X(v) for v in trace_local], [X(v) for v in trace_next]
    pl, pn = unflatten(perm_local), unflatten(perm_next)
    parts = unflatten(quotient_chunks)
    alpha_x, folded = X(alpha), X.of(0)
    r1, r2 = X(ch15[5:10]), X(ch15[10:15])
    alphas_global, acc = [], X.of(1)
    for _ in range(4):
        acc = acc * r1
        alphas_global.append(acc)

    def assert_zero(c):
        nonlocal folded
        folded = folded * alpha_x + c

    if AIRS[chip]:
        for c in AIRS[chip](loc, nxt, sel):
            assert_zero(c)
    inter = CHIPS[chip]
    k = len(inter)
    field = lambda col, row: (row[col[1]] if col[0] == "main" else X.of(col[1]) if col[0] == "const"
                              else sum(row[c] * w for c, w in col[1]) if col[0] == "weighted" else sum(row[c] for c in col[1]))
    lhs, rhs, phi0 = pn[k] - pl[k], X.of(0), X.of(0)
    for m, (sign, bus, fields, count) in enumerate(inter):
        rlc, beta = X.of(0), X.of(1)
        for f in fields:
            rlc = rlc + beta * field(f, loc)
            beta = beta * r2
        rlc = rlc + alphas_global[bus]
        assert_zero(rlc * pl[m] - 1)
        t_loc, t_nxt = pl[m] * field(count, loc), pn[m] * field(count, nxt)
        if sign == SEND:
            phi0, rhs = phi0 + t_loc, rhs + t_nxt
        else:
            phi0, rhs = phi0 - t_loc, rhs - t_nxt
    assert_zero((lhs - rhs) * sel["transition"])
    assert_zero((pl[k] - phi0) * sel["first"])
    assert_zero((pl[k] - X(cumsum)) * sel["last"])
    # reverse_slice_index_bits(&mut quotient_parts) is the identity on two parts; quotient = sum_i zeta^i * part_i
    quotient = parts[0] + zeta * parts[1]
    assert len(parts) == 2
    return folded, z_h * quotient


@pytest.mark.parametrize("workload", ["fib3", "static_data", "config5"])
def test_in_repo_verifier_identity_holds_on_oracle_proofs(built, oracle, workload):
    """folded_constraints(zeta) == Z_H(zeta) * quotient(zeta) for every chip of a real proof, with the constraints evaluated by the
    Python transcriptions on the OPENED values (VerifierConstraintFolder semantics) and the quotient recombined as verify.rs does."""
    import programs
    import valida_b200 as vb

    if workload == "fib3":
        t = vb.run_program(vb.fib_program(3), initial_fp=0x1000)
    elif workload == "static_data":
        prog, cells = programs.static_data_program()
        t = vb.run_program(prog, initial_fp=0x1000, static_data=cells)
    else:
        t = vb.run_program(programs.config5_program(6), initial_fp=0x1000)
    pr = oracle.prove(t.main, t.preprocessed, debug_checks=False)
    tr = pr.transcript()
    rows = lambda a: [[int(v) for v in r] for r in a]
    for chip in range(14):
        log_degree = t.main[chip].shape[0].bit_length() - 1
        ov = [rows(pr.opened(chip, w)) for w in range(5)]
        folded, rhs = verify_constraints_py(chip, log_degree, ov[0], ov[1], ov[2], ov[3], ov[4], [int(v) for v in pr.cumulative_sum(chip)],
                                            [int(v) for v in tr["zeta"]], [int(v) for v in tr["alpha"]], [int(v) for v in tr["perm_challenges"]])
        assert folded == rhs, (workload, chip)


# ---- get_log_quotient_degree (machine/src/symbolic/symbolic_builder.rs:17-43, symbolic_expression.rs:41-61) ------------------------
class Deg:
    """degree_multiple of a symbolic expression: variables 1, is_first_row / is_last_row 1, is_transition 0, constants 0; a sum takes
    the maximum, a product the sum."""
    __slots__ = ("d",)

    def __init__(self, d): self.d = d
    @staticmethod
    def of(v): return v if isinstance(v, Deg) else Deg(0)
    def __add__(self, o): return Deg(max(self.d, Deg.of(o).d))
    __radd__ = __sub__ = __rsub__ = __add__
    def __mul__(self, o): return Deg(self.d + Deg.of(o).d)
    __rmul__ = __mul__
    def __mod__(self, m): return self


def test_every_air_has_constraint_degree_at_most_three_so_the_quotient_has_two_chunks(built, oracle):
    """log_quotient_degree = log2_ceil(max(max_constraint_degree, 3) - 1); only Air::eval counts (the permutation constraints are not
    part of get_symbolic_constraints).  Degree <= 3 on every chip => 1 => two quotient chunks, as the oracle and the kernels assume."""
    sel = {"first": Deg(1), "last": Deg(1), "transition": Deg(0)}
    for chip, air in AIRS.items():
        if air is None:
            continue
        w = oracle.chip_width(chip)
        degs = [Deg.of(c).d for c in air([Deg(1)] * w, [Deg(1)] * w, sel)]
        constraint_degree = max(max(degs), 3)
        assert (constraint_degree - 1 - 1).bit_length() == 1, (chip, max(degs))        # log2_ceil(constraint_degree - 1)
        assert max(degs) <= 3, (chip, degs)
