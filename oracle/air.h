// ORACLE — TEST INFRASTRUCTURE ONLY (see field.h header).
// Constraint folders and the 14 BasicMachine AIRs + bus interactions, restated from the reference:
//   folders        machine/src/folding_builder.rs:32-125, machine/src/debug_builder.rs:48-54
//   AirBuilder     p3-air: when/when_ne/when_transition/assert_eq/assert_bool lower to
//                  assert_zero(filter * x)  [P3-UNVERIFIED; SURVEY App. A item 18]
//   chips          cited per function below (chip order: basic/src/lib.rs:151-166)
// Written as templates over the expression type E so that ONE transcription serves the debug
// check on the trace (E = Fp), the quotient on the LDE coset (E = Fp, ext accumulator) and the
// verifier at zeta (E = Ext5) — exactly the three builders the reference instantiates.
#pragma once
#include "field.h"
#include <functional>

namespace orc {

struct Fp {
    uint32_t v;
    Fp() : v(0) {}
    explicit Fp(uint32_t x) : v(x) {}
};
static inline Fp operator+(Fp a, Fp b) { return Fp(add(a.v, b.v)); }
static inline Fp operator-(Fp a, Fp b) { return Fp(sub(a.v, b.v)); }
static inline Fp operator*(Fp a, Fp b) { return Fp(mul(a.v, b.v)); }
static inline Fp operator-(Fp a) { return Fp(neg(a.v)); }

template <class E> struct Lift;
template <> struct Lift<Fp> {
    static Fp konst(uint32_t c) { return Fp(c % P); }
    static Ext5 to_ext(Fp x) { return Ext5::from_base(x.v); }
    static Ext5 emul(const Ext5& a, Fp b) { return a * b.v; }
    static bool is_zero(Fp x) { return x.v == 0; }
};
template <> struct Lift<Ext5> {
    static Ext5 konst(uint32_t c) { return Ext5::from_base(c % P); }
    static Ext5 to_ext(const Ext5& x) { return x; }
    static Ext5 emul(const Ext5& a, const Ext5& b) { return a * b; }
    static bool is_zero(const Ext5& x) { return x.is_zero(); }
};

// ProverConstraintFolder / VerifierConstraintFolder / DebugConstraintBuilder in one.
template <class E>
struct Folder {
    const E* main_local = nullptr; const E* main_next = nullptr;
    const E* prep_local = nullptr; const E* prep_next = nullptr;
    const Ext5* perm_local = nullptr; const Ext5* perm_next = nullptr;
    size_t perm_width = 0;
    const Ext5* perm_challenges = nullptr;  // 3
    E is_first_row, is_last_row, is_transition;
    Ext5 alpha, accumulator = Ext5::zero();
    bool debug = false;       // DebugConstraintBuilder: every constraint must vanish
    int n_constraints = 0;
    int first_failed = -1;

    static E c(uint32_t x) { return Lift<E>::konst(x); }
    E one() const { return c(1); }
    E zero() const { return c(0); }

    void assert_zero(const E& x) {
        if (debug) { if (!Lift<E>::is_zero(x) && first_failed < 0) first_failed = n_constraints; }
        else accumulator = accumulator * alpha + Lift<E>::to_ext(x);
        n_constraints++;
    }
    void assert_zero_ext(const Ext5& x) {
        if (debug) { if (!x.is_zero() && first_failed < 0) first_failed = n_constraints; }
        else accumulator = accumulator * alpha + x;
        n_constraints++;
    }
    void assert_eq(const E& x, const E& y) { assert_zero(x - y); }
    void assert_one(const E& x) { assert_zero(x - one()); }
    void assert_bool(const E& x) { assert_zero(x * (x - one())); }

    struct Filtered {
        Folder& b; E cond;
        Filtered when(const E& c2) { return Filtered{b, cond * c2}; }
        Filtered when_ne(const E& x, const E& y) { return when(x - y); }
        void assert_zero(const E& x) { b.assert_zero(cond * x); }
        void assert_eq(const E& x, const E& y) { assert_zero(x - y); }
        void assert_one(const E& x) { assert_zero(x - b.one()); }
        void assert_zero_ext(const Ext5& x) { b.assert_zero_ext(Lift<E>::emul(x, cond)); }
        void assert_eq_ext(const Ext5& x, const Ext5& y) { assert_zero_ext(x - y); }
    };
    Filtered when(const E& cnd) { return Filtered{*this, cnd}; }
    Filtered when_ne(const E& x, const E& y) { return when(x - y); }
    Filtered when_first_row() { return when(is_first_row); }
    Filtered when_last_row() { return when(is_last_row); }
    Filtered when_transition() { return when(is_transition); }
};

// ---------------------------------------------------------------------------------------------
// Bus interactions: p3_air::VirtualPairCol + machine/src/chip.rs:76-80
struct VirtualPairCol {
    struct Term { bool prep; uint32_t col; uint32_t weight; };
    std::vector<Term> terms;
    uint32_t constant = 0;
    static VirtualPairCol single_main(uint32_t c) { VirtualPairCol v; v.terms.push_back({false, c, 1}); return v; }
    static VirtualPairCol konst(uint32_t x) { VirtualPairCol v; v.constant = x; return v; }
    static VirtualPairCol sum_main(std::initializer_list<uint32_t> cols) { VirtualPairCol v; for (auto c : cols) v.terms.push_back({false, c, 1}); return v; }
    static VirtualPairCol new_main(std::initializer_list<std::pair<uint32_t, uint32_t>> cw, uint32_t k) { VirtualPairCol v; for (auto& p : cw) v.terms.push_back({false, p.first, p.second}); v.constant = k; return v; }
    template <class E> E apply(const E* prep, const E* main) const {
        E r = Lift<E>::konst(constant);
        for (auto& t : terms) r = r + (t.prep ? prep[t.col] : main[t.col]) * Lift<E>::konst(t.weight);
        return r;
    }
};
struct Interaction {
    std::vector<VirtualPairCol> fields;
    VirtualPairCol count;
    uint32_t bus;      // BusArgument::Global(bus); no chip of BasicMachine returns local interactions
    bool is_send;
};
enum : uint32_t { BUS_GENERAL = 0, BUS_PROGRAM = 1, BUS_MEM = 2, BUS_RANGE = 3 };  // basic/src/lib.rs:1190-1212

// opcodes/src/lib.rs:5-45
enum : uint32_t { OP_LOAD32 = 1, OP_STORE32 = 2, OP_JAL = 3, OP_JALV = 4, OP_BEQ = 5, OP_BNE = 6, OP_IMM32 = 7, OP_STOP = 8,
                  OP_READ_ADVICE = 9, OP_LOADFP = 10, OP_ADD32 = 100, OP_SUB32 = 101, OP_MUL32 = 102, OP_DIV32 = 103, OP_SDIV32 = 110,
                  OP_LT32 = 104, OP_SHL32 = 105, OP_SHR32 = 106, OP_AND32 = 107, OP_OR32 = 108, OP_XOR32 = 109, OP_NE32 = 111,
                  OP_MULHU32 = 112, OP_SRA32 = 113, OP_MULHS32 = 114, OP_LTE32 = 115, OP_EQ32 = 116, OP_SLT32 = 117, OP_SLE32 = 118,
                  OP_WRITE = 300 };
constexpr uint32_t BYTES_PER_INSTR = 24;

// ---- column maps ----------------------------------------------------------------------------
namespace cpu_col {  // cpu/src/columns.rs:8-37 (51 columns)
enum : uint32_t { CLK = 0, PC = 1, FP = 2, OPCODE = 3, OP_A = 4, OP_B = 5, OP_C = 6, OP_D = 7, OP_E = 8,
                  IS_BUS_OP = 9, IS_BUS_OP_WITH_MEM = 10, IS_IMM_OP = 11, IS_LEFT_IMM_OP = 12, IS_LOAD = 13, IS_LOAD_U8 = 14, IS_LOAD_S8 = 15,
                  IS_STORE = 16, IS_STORE_U8 = 17, IS_BEQ = 18, IS_BNE = 19, IS_JAL = 20, IS_JALV = 21, IS_IMM32 = 22, IS_ADVICE = 23,
                  IS_STOP = 24, IS_LOADFP = 25, DIFF = 26, DIFF_INV = 27, NOT_EQUAL = 28,
                  MEM0 = 29 /* used,is_read,addr,value[4] */, MEM1 = 36, MEM2 = 43, CLK_OR_ZERO = 50, WIDTH = 51 };
enum : uint32_t { CH_USED = 0, CH_IS_READ = 1, CH_ADDR = 2, CH_VALUE = 3 };
}
namespace mem_col {  // memory/src/columns.rs:8-39 (14)
enum : uint32_t { ADDR = 0, VALUE = 1, CLK = 5, IS_STATIC_INITIAL = 6, IS_READ = 7, IS_WRITE = 8, DIFF = 9, DIFF_INV = 10, ADDR_NOT_EQUAL = 11, COUNTER = 12, COUNTER_MULT = 13, WIDTH = 14 };
}
namespace add_col { enum : uint32_t { IN1 = 0, IN2 = 4, CARRY = 8, OUT = 11, IS_REAL = 15, WIDTH = 16 }; }   // alu_u32/src/add/columns.rs:8-18
namespace sub_col { enum : uint32_t { IN1 = 0, IN2 = 4, BORROW = 8, OUT = 11, IS_REAL = 15, WIDTH = 16 }; }  // alu_u32/src/sub/columns.rs
namespace mul_col { enum : uint32_t { IN1 = 0, IN2 = 4, OUT = 8, R = 12, S = 13, IS_MUL = 14, IS_MULHS = 15, IS_MULHU = 16, COUNTER = 17, WIDTH = 18 }; }
namespace div_col { enum : uint32_t { IN1 = 0, IN2 = 4, OUT = 8, IS_DIV = 12, IS_SDIV = 13, WIDTH = 14 }; }
namespace shift_col { enum : uint32_t { IN1 = 0, IN2 = 4, OUT = 8, BITS2 = 12, TEMP1 = 20, POW2 = 21, IS_SHL = 25, IS_SHR = 26, IS_SRA = 27, WIDTH = 28 }; }
namespace lt_col { enum : uint32_t { IN1 = 0, IN2 = 4, BYTE_FLAG = 8, BITS = 12, OUTPUT = 21, MULT = 22, IS_LT = 23, IS_LTE = 24, IS_SLT = 25, IS_SLE = 26,
                                    DIFF_INV = 27, TOP1 = 28, TOP2 = 36, DIFFERENT_SIGNS = 44, WIDTH = 45 }; }
namespace com_col { enum : uint32_t { IN1 = 0, IN2 = 4, DIFF = 8, DIFF_INV = 9, NOT_EQUAL = 10, OUTPUT = 11, IS_NE = 12, IS_EQ = 13, WIDTH = 14 }; }
namespace bw_col { enum : uint32_t { IN1 = 0, IN2 = 4, BITS1 = 8, BITS2 = 40, OUT = 72, IS_AND = 76, IS_OR = 77, IS_XOR = 78, WIDTH = 79 }; }
namespace out_col { enum : uint32_t { CLK = 0, VALUE = 1, IS_REAL = 2, DIFF = 3, COUNTER = 4, COUNTER_MULT = 5, OPCODE = 6, WIDTH = 7 }; }
namespace range_col { enum : uint32_t { MULT = 0, COUNTER = 1, WIDTH = 2 }; }
namespace sd_col { enum : uint32_t { ADDR = 0, VALUE = 1, IS_REAL = 5, WIDTH = 6 }; }

// ---- AIRs -----------------------------------------------------------------------------------
template <class E> static E reduce_word(const E* base4, const E* w) { return base4[0] * w[0] + base4[1] * w[1] + base4[2] * w[2] + base4[3] * w[3]; }
template <class E> static E sq_diff_sum(const E* a, const E* b) {
    E s = (a[0] - b[0]) * (a[0] - b[0]);
    for (int i = 1; i < 4; i++) s = s + (a[i] - b[i]) * (a[i] - b[i]);
    return s;
}

// cpu/src/stark.rs:22-305
template <class E> void eval_cpu(Folder<E>& b) {
    using namespace cpu_col;
    const E* l = b.main_local; const E* n = b.main_next;
    E one = b.one();
    E base[4] = {b.c(1u << 24), b.c(1u << 16), b.c(1u << 8), b.c(1)};
    E bpi = b.c(BYTES_PER_INSTR);
    const E* rv1 = l + MEM0 + CH_VALUE; const E* rv2 = l + MEM1 + CH_VALUE; const E* wv = l + MEM2 + CH_VALUE;
    E ra1 = l[MEM0 + CH_ADDR], ra2 = l[MEM1 + CH_ADDR], wa = l[MEM2 + CH_ADDR];
    E r1u = l[MEM0 + CH_USED], r2u = l[MEM1 + CH_USED], wu = l[MEM2 + CH_USED];
    // eval_pc (207-258)
    {
        E should_inc = l[IS_IMM32] + l[IS_LOADFP] + l[IS_BUS_OP] + l[IS_ADVICE];
        E inc_pc = l[PC] + one;
        b.when_transition().when(should_inc).assert_eq(n[PC], inc_pc);
        E equal = one - l[NOT_EQUAL];
        E npc24 = l[OP_A];
        E beq24 = equal * npc24 + bpi * l[NOT_EQUAL] * inc_pc;
        E bne24 = bpi * equal * inc_pc + l[NOT_EQUAL] * npc24;
        b.when_transition().when(l[IS_BEQ]).assert_eq(bpi * n[PC], beq24);
        b.when_transition().when(l[IS_BNE]).assert_eq(bpi * n[PC], bne24);
        b.when_transition().when(l[IS_JAL]).assert_eq(bpi * n[PC], l[OP_B]);
        b.when_transition().when(l[IS_JALV]).assert_eq(bpi * n[PC], reduce_word(base, rv1));
    }
    // eval_fp (260-281)
    b.when_transition().when(l[IS_JAL]).assert_eq(n[FP], l[FP] + l[OP_C]);
    b.when_transition().when(l[IS_JALV]).assert_eq(n[FP], l[FP] + reduce_word(base, rv2));
    b.when_transition().when(one - l[IS_JAL] - l[IS_JALV]).assert_eq(n[FP], l[FP]);
    // eval_equality (283-305)
    b.assert_eq(l[DIFF], sq_diff_sum(rv1, rv2));
    b.assert_bool(l[NOT_EQUAL]);
    b.assert_eq(l[NOT_EQUAL], l[DIFF] * l[DIFF_INV]);
    b.assert_zero((one - l[NOT_EQUAL]) * l[DIFF]);
    // eval_memory_channels (71-205)
    {
        E is_load = l[IS_LOAD], is_store = l[IS_STORE], is_jal = l[IS_JAL], is_jalv = l[IS_JALV], is_beq = l[IS_BEQ], is_bne = l[IS_BNE],
          is_imm32 = l[IS_IMM32], is_loadfp = l[IS_LOADFP], is_imm_op = l[IS_IMM_OP], is_left_imm_op = l[IS_LEFT_IMM_OP], is_bus_op = l[IS_BUS_OP];
        b.assert_bool(is_load); b.assert_bool(is_store); b.assert_bool(is_jal); b.assert_bool(is_jalv); b.assert_bool(is_beq); b.assert_bool(is_bne);
        b.assert_bool(is_imm32); b.assert_bool(is_loadfp); b.assert_bool(is_imm_op); b.assert_bool(is_left_imm_op); b.assert_bool(is_bus_op);
        E addr_a = l[FP] + l[OP_A], addr_b = l[FP] + l[OP_B], addr_c = l[FP] + l[OP_C];
        b.assert_one(l[MEM0 + CH_IS_READ]);
        b.assert_one(l[MEM1 + CH_IS_READ]);
        b.assert_zero(l[MEM2 + CH_IS_READ]);
        // Read (1)
        b.when(is_jalv + is_beq + is_bne + is_bus_op * (one - is_left_imm_op)).assert_eq(ra1, addr_b);
        b.when(is_load + is_store).assert_eq(ra1, addr_c);
        b.when(is_load + is_store + is_jalv + is_beq + is_bne + (one - is_left_imm_op) * is_bus_op).assert_one(r1u);
        b.when(is_jal + is_left_imm_op + is_loadfp + is_imm32).assert_zero(r1u);
        // Read (2)
        b.when(is_load).assert_eq(ra2, reduce_word(base, rv1));
        b.when(is_store).assert_eq(ra2, addr_b);
        b.when(is_jalv + (one - is_imm_op) * is_bus_op).assert_eq(ra2, addr_c);
        b.when(is_load + is_store + is_jalv + (one - is_imm_op) * (is_beq + is_bne + is_bus_op)).assert_one(r2u);
        b.when(is_jal + is_imm_op * (is_beq + is_bne + is_bus_op) + is_loadfp + is_imm32).assert_zero(r2u);
        // Write
        b.when(is_load + is_jal + is_jalv + is_imm32 + is_bus_op + is_loadfp).assert_eq(wa, addr_a);
        b.when(is_store).assert_eq(wa, reduce_word(base, rv2));
        b.when(is_store).assert_zero(sq_diff_sum(rv1, wv));
        b.when(is_load).assert_zero(sq_diff_sum(rv2, wv));
        b.when_transition().when(is_jal + is_jalv).assert_eq(bpi * (l[PC] + one), reduce_word(base, wv));
        b.when(is_imm32).assert_zero(sq_diff_sum(wv, l + OP_B));
        b.when(is_loadfp).assert_eq(addr_b, reduce_word(base, wv));
        b.when(is_store + is_load + is_jal + is_jalv + is_imm32 + is_loadfp + is_bus_op).assert_one(wu);
        b.when(is_beq + is_bne).assert_zero(wu);
    }
    // Clock constraints (39-49)
    b.when_first_row().assert_zero(l[CLK]);
    b.when_transition().assert_eq(l[CLK] + one, n[CLK]);
    b.when(l[IS_BUS_OP_WITH_MEM]).assert_eq(l[CLK], l[CLK_OR_ZERO]);
    b.when(one - l[IS_BUS_OP_WITH_MEM]).assert_zero(l[CLK_OR_ZERO]);
    // Immediate value constraints (51-62)
    b.assert_bool(l[IS_IMM_OP] + l[IS_LEFT_IMM_OP]);
    b.when(l[IS_IMM_OP]).assert_eq(l[OP_C], reduce_word(base, rv2));
    b.when(l[IS_LEFT_IMM_OP]).assert_eq(l[OP_B], reduce_word(base, rv1));
    // Stop constraints (64-71)
    b.when_transition().when(l[IS_STOP]).assert_eq(n[PC], l[PC]);
    b.when_last_row().assert_one(l[IS_STOP]);
}

// alu_u32/src/add/stark.rs:21-55
template <class E> void eval_add(Folder<E>& b) {
    using namespace add_col;
    const E* l = b.main_local;
    E one = b.one(), base = b.c(1u << 8);
    E c1 = l[CARRY], c2 = l[CARRY + 1], c3 = l[CARRY + 2];
    E o0 = l[IN1 + 3] + l[IN2 + 3] - l[OUT + 3];
    E o1 = l[IN1 + 2] + l[IN2 + 2] - l[OUT + 2] + c1;
    E o2 = l[IN1 + 1] + l[IN2 + 1] - l[OUT + 1] + c2;
    E o3 = l[IN1 + 0] + l[IN2 + 0] - l[OUT + 0] + c3;
    b.assert_zero(o0 * (o0 - base));
    b.assert_zero(o1 * (o1 - base));
    b.assert_zero(o2 * (o2 - base));
    b.assert_zero(o3 * (o3 - base));
    b.assert_zero(o0 * (c1 - one) + (o0 - base) * c1);
    b.assert_zero(o1 * (c2 - one) + (o1 - base) * c2);
    b.assert_zero(o2 * (c3 - one) + (o2 - base) * c3);
    b.assert_bool(c1); b.assert_bool(c2); b.assert_bool(c3);
}

// alu_u32/src/sub/stark.rs:21-52
template <class E> void eval_sub(Folder<E>& b) {
    using namespace sub_col;
    const E* l = b.main_local;
    E base = b.c(1u << 8);
    E b1 = l[BORROW], b2 = l[BORROW + 1], b3 = l[BORROW + 2];
    b.assert_eq(l[OUT + 3], base * b1 + l[IN1 + 3] - l[IN2 + 3]);
    b.assert_eq(l[OUT + 2], base * b2 + l[IN1 + 2] - l[IN2 + 2] - b1);
    b.assert_eq(l[OUT + 1], base * b3 + l[IN1 + 1] - l[IN2 + 1] - b2);
    b.assert_eq(l[OUT + 0], l[IN1 + 0] - l[IN2 + 0] - b3);
    b.assert_bool(b1); b.assert_bool(b2); b.assert_bool(b3);
}

// alu_u32/src/mul/stark.rs:23-82
template <class E> void eval_mul(Folder<E>& b) {
    using namespace mul_col;
    const E* l = b.main_local; const E* n = b.main_next;
    E base_m[4] = {b.c(1), b.c(1u << 8), b.c(1u << 16), b.c(1u << 24)};
    auto pi_m = [&](int N) { E s = b.zero(); for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) if (i + j < N) s = s + base_m[i + j] * l[IN1 + 3 - i] * l[IN2 + 3 - j]; return s; };
    auto sigma_m = [&](int N) { E s = b.zero(); for (int i = 0; i < N; i++) s = s + base_m[i] * l[OUT + 3 - i]; return s; };
    E pi = pi_m(4), sigma = sigma_m(4), pi_p = pi_m(2), sigma_p = sigma_m(2);
    b.assert_eq(pi - sigma, l[R] * b.c(2));
    b.assert_eq(pi_p - sigma_p, l[S] * base_m[2]);
    b.when_first_row().assert_eq(l[COUNTER], b.one());
    E cd = n[COUNTER] - l[COUNTER];
    b.when_transition().assert_zero(cd * (cd - b.one()));
    b.when_last_row().assert_eq(l[COUNTER], b.c(1u << 10));
}

// alu_u32/src/shift/stark.rs:21-70
template <class E> void eval_shift(Folder<E>& b) {
    using namespace shift_col;
    const E* l = b.main_local;
    E one = b.one();
    E byte2 = b.zero();
    for (int i = 0; i < 8; i++) byte2 = byte2 + l[BITS2 + i] * b.c(1u << i);
    b.assert_eq(l[IN2 + 3], byte2);
    for (int i = 0; i < 8; i++) b.assert_bool(l[BITS2 + i]);
    E temp1 = (l[BITS2 + 0] * b.c(1u << 1)) * (l[BITS2 + 1] * b.c(1u << 2)) * (l[BITS2 + 2] * b.c(1u << 4));
    b.assert_eq(l[TEMP1], temp1);
    b.assert_eq(l[POW2 + 0], l[TEMP1] * (one - l[BITS2 + 3]) * (one - l[BITS2 + 4]));
    b.assert_eq(l[POW2 + 1], l[TEMP1] * l[BITS2 + 3] * (one - l[BITS2 + 4]));
    b.assert_eq(l[POW2 + 2], l[TEMP1] * (one - l[BITS2 + 3]) * l[BITS2 + 4]);
    b.assert_eq(l[POW2 + 3], l[TEMP1] * l[BITS2 + 3] * l[BITS2 + 4]);
    b.assert_bool(l[IS_SHL]); b.assert_bool(l[IS_SHR]); b.assert_bool(l[IS_SRA]);
    b.assert_bool(l[IS_SHL] + l[IS_SHR] + l[IS_SRA]);
}

// alu_u32/src/lt/stark.rs:21-169
template <class E> void eval_lt(Folder<E>& b) {
    using namespace lt_col;
    const E* l = b.main_local;
    E one = b.one();
    E bit_comp = b.zero();
    for (int i = 0; i < 9; i++) bit_comp = bit_comp + l[BITS + i] * b.c(1u << i);
    E flag_sum = l[BYTE_FLAG] + l[BYTE_FLAG + 1] + l[BYTE_FLAG + 2] + l[BYTE_FLAG + 3];
    b.assert_bool(flag_sum);
    b.when_ne(l[BYTE_FLAG], one).assert_eq(l[IN1 + 0], l[IN2 + 0]);
    b.when_ne(l[BYTE_FLAG] + l[BYTE_FLAG + 1], one).assert_eq(l[IN1 + 1], l[IN2 + 1]);
    b.when_ne(l[BYTE_FLAG] + l[BYTE_FLAG + 1] + l[BYTE_FLAG + 2], one).assert_eq(l[IN1 + 2], l[IN2 + 2]);
    b.when_ne(flag_sum, one).assert_eq(l[IN1 + 3], l[IN2 + 3]);
    b.when_ne(flag_sum, one).assert_eq(bit_comp, b.zero());
    for (int i = 0; i < 4; i++) {
        b.when(l[BYTE_FLAG + i]).assert_eq(b.c(256) + l[IN1 + i] - l[IN2 + i], bit_comp);
        b.when(l[BYTE_FLAG + i]).assert_eq((l[IN1 + i] - l[IN2 + i]) * l[DIFF_INV], one);
        b.assert_bool(l[BYTE_FLAG + i]);
    }
    E top1 = b.zero(), top2 = b.zero();
    for (int i = 0; i < 8; i++) { top1 = top1 + l[TOP1 + i] * b.c(1u << i); top2 = top2 + l[TOP2 + i] * b.c(1u << i); }
    b.assert_eq(top1, l[IN1 + 0]);
    b.assert_eq(top2, l[IN2 + 0]);
    E is_signed = l[IS_SLT] + l[IS_SLE];
    E is_unsigned = one - is_signed;
    E same_sign = one - l[DIFFERENT_SIGNS];
    E are_equal = one - flag_sum;
    b.when(is_unsigned).assert_zero(l[DIFFERENT_SIGNS]);
    b.when(is_signed).when_ne(l[TOP1 + 7], l[TOP2 + 7]).assert_eq(l[DIFFERENT_SIGNS], one);
    b.when(l[DIFFERENT_SIGNS]).assert_eq(l[BYTE_FLAG], one);
    b.when(l[DIFFERENT_SIGNS]).assert_eq(l[TOP1 + 7] + l[TOP2 + 7], one);
    b.assert_bool(l[IS_LT]); b.assert_bool(l[IS_LTE]); b.assert_bool(l[IS_SLT]); b.assert_bool(l[IS_SLE]);
    b.assert_bool(l[IS_LT] + l[IS_LTE] + l[IS_SLT] + l[IS_SLE]);
    b.when(l[BITS + 8]).when(is_unsigned + same_sign).assert_zero(l[OUTPUT]);
    b.when(l[BITS + 8]).when(l[DIFFERENT_SIGNS]).assert_one(l[OUTPUT]);
    b.when_ne(l[BITS + 8] + are_equal, one).when(is_unsigned + same_sign).assert_one(l[OUTPUT]);
    b.when_ne(l[BITS + 8] + are_equal, one).when(l[DIFFERENT_SIGNS]).assert_zero(l[OUTPUT]);
    b.when(are_equal).when(l[IS_LTE] + l[IS_SLE]).assert_one(l[OUTPUT]);
    b.when(are_equal).when(l[IS_LT] + l[IS_SLT]).assert_zero(l[OUTPUT]);
    for (int i = 0; i < 9; i++) b.assert_bool(l[BITS + i]);
    for (int i = 0; i < 8; i++) b.assert_bool(l[TOP1 + i]);
    for (int i = 0; i < 8; i++) b.assert_bool(l[TOP2 + i]);
}

// alu_u32/src/com/stark.rs:21-50
template <class E> void eval_com(Folder<E>& b) {
    using namespace com_col;
    const E* l = b.main_local;
    E one = b.one();
    b.assert_eq(l[DIFF], sq_diff_sum(l + IN1, l + IN2));
    b.assert_bool(l[NOT_EQUAL]);
    b.assert_eq(l[NOT_EQUAL], l[DIFF] * l[DIFF_INV]);
    b.assert_zero((one - l[NOT_EQUAL]) * l[DIFF]);
    b.assert_bool(l[IS_NE]); b.assert_bool(l[IS_EQ]); b.assert_bool(l[IS_NE] + l[IS_EQ]);
    b.assert_eq(l[OUTPUT], l[IS_NE] * l[NOT_EQUAL] + l[IS_EQ] * (one - l[NOT_EQUAL]));
}

// alu_u32/src/bitwise/stark.rs:22-75
template <class E> void eval_bitwise(Folder<E>& b) {
    using namespace bw_col;
    const E* l = b.main_local;
    for (int i = 0; i < 4; i++) {
        E byte1 = b.zero(), byte2 = b.zero(), band = b.zero();
        for (int k = 0; k < 8; k++) {
            byte1 = byte1 + l[BITS1 + 8 * i + k] * b.c(1u << k);
            byte2 = byte2 + l[BITS2 + 8 * i + k] * b.c(1u << k);
            band = band + l[BITS1 + 8 * i + k] * l[BITS2 + 8 * i + k] * b.c(1u << k);
        }
        b.assert_eq(l[IN1 + i], byte1);
        b.assert_eq(l[IN2 + i], byte2);
        E bor = byte1 + byte2 - band;
        E bxor = byte1 + byte2 - b.c(2) * band;
        b.when(l[IS_AND]).assert_eq(band, l[OUT + i]);
        b.when(l[IS_OR]).assert_eq(bor, l[OUT + i]);
        b.when(l[IS_XOR]).assert_eq(bxor, l[OUT + i]);
        for (int k = 0; k < 8; k++) b.assert_bool(l[BITS1 + 8 * i + k]);
        for (int k = 0; k < 8; k++) b.assert_bool(l[BITS2 + 8 * i + k]);
    }
    b.assert_bool(l[IS_AND]); b.assert_bool(l[IS_OR]); b.assert_bool(l[IS_XOR]);
    b.assert_bool(l[IS_AND] + l[IS_OR] + l[IS_XOR]);
}

// output/src/stark.rs:21-39
template <class E> void eval_output(Folder<E>& b) {
    using namespace out_col;
    const E* l = b.main_local; const E* n = b.main_next;
    b.when_transition().assert_eq(l[DIFF], n[CLK] - l[CLK]);
    b.when_transition().assert_eq(n[COUNTER], l[COUNTER] + b.one());
    b.when(l[IS_REAL]).assert_eq(l[OPCODE], b.c(OP_WRITE));
}

// static_data/src/stark.rs:25-37
template <class E> void eval_static_data(Folder<E>& b) {
    using namespace sd_col;
    const E* l = b.main_local; const E* n = b.main_next;
    b.when_transition().when(l[IS_REAL] * n[IS_REAL]).assert_eq(n[ADDR], l[ADDR] + b.one() + b.one() + b.one() + b.one());
}

// memory/src/stark.rs:22-79, range/src/stark.rs:12-14, program/src/stark.rs:14, alu_u32/src/div/stark.rs:18-20: empty
template <class E> void eval_empty(Folder<E>&) {}

// ---- chip table -------------------------------------------------------------------------------
struct ChipDef {
    const char* name;
    uint32_t width, prep_width;
    void (*eval_fp)(Folder<Fp>&);
    void (*eval_ext)(Folder<Ext5>&);
    std::vector<Interaction> interactions;  // all_interactions order: sends then receives (machine/src/chip.rs:40-63)
};
constexpr int NUM_CHIPS = 14;
enum { CHIP_CPU = 0, CHIP_PROGRAM, CHIP_MEM, CHIP_ADD, CHIP_SUB, CHIP_MUL, CHIP_DIV, CHIP_SHIFT, CHIP_LT, CHIP_COM, CHIP_BITWISE, CHIP_OUTPUT, CHIP_RANGE, CHIP_STATIC };

static inline std::vector<VirtualPairCol> word_cols(uint32_t start) {
    std::vector<VirtualPairCol> v;
    for (uint32_t i = 0; i < 4; i++) v.push_back(VirtualPairCol::single_main(start + i));
    return v;
}
static inline void extend(std::vector<VirtualPairCol>& a, const std::vector<VirtualPairCol>& b) { a.insert(a.end(), b.begin(), b.end()); }
static inline Interaction alu_bus(VirtualPairCol opcode, uint32_t in1, uint32_t in2, uint32_t out, VirtualPairCol count, bool is_send) {
    Interaction it; it.fields.push_back(opcode); extend(it.fields, word_cols(in1)); extend(it.fields, word_cols(in2)); extend(it.fields, word_cols(out));
    it.count = count; it.bus = BUS_GENERAL; it.is_send = is_send; return it;
}
static inline Interaction alu_bus_scalar_out(VirtualPairCol opcode, uint32_t in1, uint32_t in2, uint32_t out, VirtualPairCol count) {
    Interaction it; it.fields.push_back(opcode); extend(it.fields, word_cols(in1)); extend(it.fields, word_cols(in2));
    for (int i = 0; i < 3; i++) it.fields.push_back(VirtualPairCol::konst(0));
    it.fields.push_back(VirtualPairCol::single_main(out));
    it.count = count; it.bus = BUS_GENERAL; it.is_send = false; return it;
}

static inline const std::vector<ChipDef>& chips() {
    static std::vector<ChipDef> defs = [] {
        std::vector<ChipDef> d(NUM_CHIPS);
        using V = VirtualPairCol;
        // Cpu: cpu/src/lib.rs:99-159 — 3 memory sends + 1 general send
        {
            ChipDef& c = d[CHIP_CPU]; c = {"cpu", cpu_col::WIDTH, 0, eval_cpu<Fp>, eval_cpu<Ext5>, {}};
            for (uint32_t ch : {cpu_col::MEM0, cpu_col::MEM1, cpu_col::MEM2}) {
                Interaction it;
                it.fields = {V::single_main(ch + cpu_col::CH_IS_READ), V::single_main(cpu_col::CLK), V::single_main(ch + cpu_col::CH_ADDR), V::konst(0)};
                extend(it.fields, word_cols(ch + cpu_col::CH_VALUE));
                it.count = V::single_main(ch + cpu_col::CH_USED); it.bus = BUS_MEM; it.is_send = true;
                c.interactions.push_back(it);
            }
            Interaction g; g.fields.push_back(V::single_main(cpu_col::OPCODE));
            for (uint32_t ch : {cpu_col::MEM0, cpu_col::MEM1, cpu_col::MEM2}) extend(g.fields, word_cols(ch + cpu_col::CH_VALUE));
            g.fields.push_back(V::single_main(cpu_col::CLK_OR_ZERO));
            g.count = V::single_main(cpu_col::IS_BUS_OP); g.bus = BUS_GENERAL; g.is_send = true;
            c.interactions.push_back(g);
        }
        // Program: program/src/lib.rs:50-68 — no interactions; 7 preprocessed columns
        d[CHIP_PROGRAM] = {"program", 1, 7, eval_empty<Fp>, eval_empty<Ext5>, {}};
        // Memory: memory/src/lib.rs:216-233
        {
            ChipDef& c = d[CHIP_MEM]; c = {"memory", mem_col::WIDTH, 0, eval_empty<Fp>, eval_empty<Ext5>, {}};
            Interaction it;
            it.fields = {V::single_main(mem_col::IS_READ), V::single_main(mem_col::CLK), V::single_main(mem_col::ADDR), V::single_main(mem_col::IS_STATIC_INITIAL)};
            extend(it.fields, word_cols(mem_col::VALUE));
            it.count = V::sum_main({mem_col::IS_READ, mem_col::IS_WRITE}); it.bus = BUS_MEM; it.is_send = false;
            c.interactions.push_back(it);
        }
        // Add32: alu_u32/src/add/mod.rs:53-88 — 4 range sends, 1 general receive
        {
            ChipDef& c = d[CHIP_ADD]; c = {"add32", add_col::WIDTH, 0, eval_add<Fp>, eval_add<Ext5>, {}};
            for (uint32_t i = 0; i < 4; i++) c.interactions.push_back({{V::single_main(add_col::OUT + i)}, V::single_main(add_col::IS_REAL), BUS_RANGE, true});
            c.interactions.push_back(alu_bus(V::konst(OP_ADD32), add_col::IN1, add_col::IN2, add_col::OUT, V::single_main(add_col::IS_REAL), false));
        }
        // Sub32: alu_u32/src/sub/mod.rs
        {
            ChipDef& c = d[CHIP_SUB]; c = {"sub32", sub_col::WIDTH, 0, eval_sub<Fp>, eval_sub<Ext5>, {}};
            for (uint32_t i = 0; i < 4; i++) c.interactions.push_back({{V::single_main(sub_col::OUT + i)}, V::single_main(sub_col::IS_REAL), BUS_RANGE, true});
            c.interactions.push_back(alu_bus(V::konst(OP_SUB32), sub_col::IN1, sub_col::IN2, sub_col::OUT, V::single_main(sub_col::IS_REAL), false));
        }
        // Mul32: alu_u32/src/mul/mod.rs:66-96
        {
            ChipDef& c = d[CHIP_MUL]; c = {"mul32", mul_col::WIDTH, 0, eval_mul<Fp>, eval_mul<Ext5>, {}};
            c.interactions.push_back(alu_bus(V::new_main({{mul_col::IS_MUL, OP_MUL32}, {mul_col::IS_MULHS, OP_MULHS32}, {mul_col::IS_MULHU, OP_MULHU32}}, 0),
                                             mul_col::IN1, mul_col::IN2, mul_col::OUT, V::sum_main({mul_col::IS_MUL, mul_col::IS_MULHS, mul_col::IS_MULHU}), false));
        }
        // Div32: alu_u32/src/div/mod.rs
        {
            ChipDef& c = d[CHIP_DIV]; c = {"div32", div_col::WIDTH, 0, eval_empty<Fp>, eval_empty<Ext5>, {}};
            c.interactions.push_back(alu_bus(V::new_main({{div_col::IS_DIV, OP_DIV32}, {div_col::IS_SDIV, OP_SDIV32}}, 0),
                                             div_col::IN1, div_col::IN2, div_col::OUT, V::sum_main({div_col::IS_DIV, div_col::IS_SDIV}), false));
        }
        // Shift32: alu_u32/src/shift/mod.rs:58-116 — 1 send (as MUL32/DIV32/SDIV32 with power_of_two), 1 receive
        {
            ChipDef& c = d[CHIP_SHIFT]; c = {"shift32", shift_col::WIDTH, 0, eval_shift<Fp>, eval_shift<Ext5>, {}};
            V real = V::sum_main({shift_col::IS_SHL, shift_col::IS_SHR, shift_col::IS_SRA});
            c.interactions.push_back(alu_bus(V::new_main({{shift_col::IS_SHL, OP_MUL32}, {shift_col::IS_SHR, OP_DIV32}, {shift_col::IS_SRA, OP_SDIV32}}, 0),
                                             shift_col::IN1, shift_col::POW2, shift_col::OUT, real, true));
            c.interactions.push_back(alu_bus(V::new_main({{shift_col::IS_SHL, OP_SHL32}, {shift_col::IS_SHR, OP_SHR32}, {shift_col::IS_SRA, OP_SRA32}}, 0),
                                             shift_col::IN1, shift_col::IN2, shift_col::OUT, real, false));
        }
        // Lt32: alu_u32/src/lt/mod.rs:58-84
        {
            ChipDef& c = d[CHIP_LT]; c = {"lt32", lt_col::WIDTH, 0, eval_lt<Fp>, eval_lt<Ext5>, {}};
            c.interactions.push_back(alu_bus_scalar_out(V::new_main({{lt_col::IS_LT, OP_LT32}, {lt_col::IS_LTE, OP_LTE32}, {lt_col::IS_SLT, OP_SLT32}, {lt_col::IS_SLE, OP_SLE32}}, 0),
                                                        lt_col::IN1, lt_col::IN2, lt_col::OUTPUT, V::single_main(lt_col::MULT)));
        }
        // Com32: alu_u32/src/com/mod.rs
        {
            ChipDef& c = d[CHIP_COM]; c = {"com32", com_col::WIDTH, 0, eval_com<Fp>, eval_com<Ext5>, {}};
            c.interactions.push_back(alu_bus_scalar_out(V::new_main({{com_col::IS_NE, OP_NE32}, {com_col::IS_EQ, OP_EQ32}}, 0),
                                                        com_col::IN1, com_col::IN2, com_col::OUTPUT, V::sum_main({com_col::IS_NE, com_col::IS_EQ})));
        }
        // Bitwise32: alu_u32/src/bitwise/mod.rs
        {
            ChipDef& c = d[CHIP_BITWISE]; c = {"bitwise32", bw_col::WIDTH, 0, eval_bitwise<Fp>, eval_bitwise<Ext5>, {}};
            c.interactions.push_back(alu_bus(V::new_main({{bw_col::IS_AND, OP_AND32}, {bw_col::IS_OR, OP_OR32}, {bw_col::IS_XOR, OP_XOR32}}, 0),
                                             bw_col::IN1, bw_col::IN2, bw_col::OUT, V::sum_main({bw_col::IS_AND, bw_col::IS_OR, bw_col::IS_XOR}), false));
        }
        // Output: output/src/lib.rs:117-136
        {
            ChipDef& c = d[CHIP_OUTPUT]; c = {"output", out_col::WIDTH, 0, eval_output<Fp>, eval_output<Ext5>, {}};
            Interaction it; it.fields.push_back(V::single_main(out_col::OPCODE));
            for (int i = 0; i < 12; i++) it.fields.push_back(i == 3 ? V::single_main(out_col::VALUE) : V::konst(0));
            it.fields.push_back(V::single_main(out_col::CLK));
            it.count = V::single_main(out_col::IS_REAL); it.bus = BUS_GENERAL; it.is_send = false;
            c.interactions.push_back(it);
        }
        // Range<256>: range/src/lib.rs:45-56 — 1 preprocessed column
        {
            ChipDef& c = d[CHIP_RANGE]; c = {"range", range_col::WIDTH, 1, eval_empty<Fp>, eval_empty<Ext5>, {}};
            c.interactions.push_back({{V::single_main(range_col::COUNTER)}, V::single_main(range_col::MULT), BUS_RANGE, false});
        }
        // StaticData: static_data/src/lib.rs:81-96
        {
            ChipDef& c = d[CHIP_STATIC]; c = {"static_data", sd_col::WIDTH, 0, eval_static_data<Fp>, eval_static_data<Ext5>, {}};
            Interaction it;
            it.fields = {V::konst(0), V::konst(0), V::single_main(sd_col::ADDR), V::konst(1)};
            extend(it.fields, word_cols(sd_col::VALUE));
            it.count = V::single_main(sd_col::IS_REAL); it.bus = BUS_MEM; it.is_send = true;
            c.interactions.push_back(it);
        }
        return d;
    }();
    return defs;
}

}  // namespace orc
