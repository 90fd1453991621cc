// BasicMachine chip table: widths and bus interactions of the 14 chips in proving order
// (basic/src/lib.rs:151-166), the data-driven mirror of each Chip::global_sends / global_receives:
//   cpu        cpu/src/lib.rs:99-159            memory     memory/src/lib.rs:216-233
//   add/sub    alu_u32/src/add/mod.rs:53-88, alu_u32/src/sub/mod.rs:53-88
//   mul        alu_u32/src/mul/mod.rs:66-96     div        alu_u32/src/div/mod.rs:53-77
//   shift      alu_u32/src/shift/mod.rs:58-116  lt         alu_u32/src/lt/mod.rs:58-84
//   com        alu_u32/src/com/mod.rs:55-82     bitwise    alu_u32/src/bitwise/mod.rs:54-81
//   output     output/src/lib.rs:117-136        range      range/src/lib.rs:45-56
//   static     static_data/src/lib.rs:81-96     program    program/src/lib.rs:50-68 (none)
// Buses: general = 0, program = 1, mem = 2, range = 3 (basic/src/lib.rs:1190-1212).
#include "../../../include/valida_b200.h"
#include <cstring>
#include <mutex>

namespace {

vgpu_chip_desc g_chips[VGPU_NUM_CHIPS];
std::once_flag g_once;

vgpu_pair_col col(uint32_t c) { vgpu_pair_col v{}; v.n_terms = 1; v.terms[0].column = c; v.terms[0].weight = 1; return v; }
vgpu_pair_col konst(uint32_t k) { vgpu_pair_col v{}; v.constant = k; return v; }
vgpu_pair_col weighted(std::initializer_list<std::pair<uint32_t, uint32_t>> cw) {
    vgpu_pair_col v{};
    for (auto& p : cw) { v.terms[v.n_terms].column = p.first; v.terms[v.n_terms].weight = p.second; v.n_terms++; }
    return v;
}
void word(vgpu_interaction& it, uint32_t first_col) { for (uint32_t i = 0; i < 4; i++) it.fields[it.n_fields++] = col(first_col + i); }
vgpu_interaction& add_interaction(vgpu_chip_desc& c, uint32_t bus, bool send, vgpu_pair_col count) {
    vgpu_interaction& it = c.interactions[c.n_interactions++];
    it.bus = bus; it.is_send = send ? 1 : 0; it.count = count;
    return it;
}
// general-bus tuple of an ALU chip: (opcode, input_1[4], input_2[4], output[4])
void alu_tuple(vgpu_interaction& it, vgpu_pair_col opcode, uint32_t in1, uint32_t in2, uint32_t out) {
    it.fields[it.n_fields++] = opcode; word(it, in1); word(it, in2); word(it, out);
}
void alu_tuple_scalar_out(vgpu_interaction& it, vgpu_pair_col opcode, uint32_t in1, uint32_t in2, uint32_t out) {
    it.fields[it.n_fields++] = opcode; word(it, in1); word(it, in2);
    for (int i = 0; i < 3; i++) it.fields[it.n_fields++] = konst(0);
    it.fields[it.n_fields++] = col(out);
}

void init() {
    std::memset(g_chips, 0, sizeof g_chips);
    const uint32_t GENERAL = 0, MEM = 2, RANGE = 3;
    auto def = [&](uint32_t id, uint32_t w, uint32_t pw) -> vgpu_chip_desc& { g_chips[id].chip_id = id; g_chips[id].width = w; g_chips[id].preprocessed_width = pw; return g_chips[id]; };
    {   // 0 cpu: mem channels at columns 29/36/43 = (used, is_read, addr, value[4]); clk = 0; opcode = 3; is_bus_op = 9; clk_or_zero = 50
        vgpu_chip_desc& c = def(0, 51, 0);
        for (uint32_t ch : {29u, 36u, 43u}) {
            vgpu_interaction& it = add_interaction(c, MEM, true, col(ch + 0));
            it.fields[it.n_fields++] = col(ch + 1); it.fields[it.n_fields++] = col(0); it.fields[it.n_fields++] = col(ch + 2); it.fields[it.n_fields++] = konst(0);
            word(it, ch + 3);
        }
        vgpu_interaction& g = add_interaction(c, GENERAL, true, col(9));
        g.fields[g.n_fields++] = col(3);
        for (uint32_t ch : {29u, 36u, 43u}) word(g, ch + 3);
        g.fields[g.n_fields++] = col(50);
    }
    def(1, 1, 7);   // program: no interactions
    {   // 2 memory: addr 0, value 1..4, clk 5, is_static_initial 6, is_read 7, is_write 8
        vgpu_chip_desc& c = def(2, 14, 0);
        vgpu_interaction& it = add_interaction(c, MEM, false, weighted({{7, 1}, {8, 1}}));
        it.fields[it.n_fields++] = col(7); it.fields[it.n_fields++] = col(5); it.fields[it.n_fields++] = col(0); it.fields[it.n_fields++] = col(6);
        word(it, 1);
    }
    for (uint32_t id : {3u, 4u}) {   // add32 / sub32: in1 0, in2 4, carry|borrow 8, out 11, is_real 15
        vgpu_chip_desc& c = def(id, 16, 0);
        for (uint32_t i = 0; i < 4; i++) { vgpu_interaction& r = add_interaction(c, RANGE, true, col(15)); r.fields[r.n_fields++] = col(11 + i); }
        vgpu_interaction& g = add_interaction(c, GENERAL, false, col(15));
        alu_tuple(g, konst(id == 3 ? 100 : 101), 0, 4, 11);
    }
    {   // 5 mul32: in1 0, in2 4, out 8, is_mul 14, is_mulhs 15, is_mulhu 16
        vgpu_chip_desc& c = def(5, 18, 0);
        vgpu_interaction& g = add_interaction(c, GENERAL, false, weighted({{14, 1}, {15, 1}, {16, 1}}));
        alu_tuple(g, weighted({{14, 102}, {15, 114}, {16, 112}}), 0, 4, 8);
    }
    {   // 6 div32: is_div 12, is_sdiv 13
        vgpu_chip_desc& c = def(6, 14, 0);
        vgpu_interaction& g = add_interaction(c, GENERAL, false, weighted({{12, 1}, {13, 1}}));
        alu_tuple(g, weighted({{12, 103}, {13, 110}}), 0, 4, 8);
    }
    {   // 7 shift32: in1 0, in2 4, out 8, power_of_two 21, is_shl 25, is_shr 26, is_sra 27
        vgpu_chip_desc& c = def(7, 28, 0);
        vgpu_pair_col real = weighted({{25, 1}, {26, 1}, {27, 1}});
        vgpu_interaction& s = add_interaction(c, GENERAL, true, real);
        alu_tuple(s, weighted({{25, 102}, {26, 103}, {27, 110}}), 0, 21, 8);
        vgpu_interaction& r = add_interaction(c, GENERAL, false, real);
        alu_tuple(r, weighted({{25, 105}, {26, 106}, {27, 113}}), 0, 4, 8);
    }
    {   // 8 lt32: output 21, multiplicity 22, is_lt 23, is_lte 24, is_slt 25, is_sle 26
        vgpu_chip_desc& c = def(8, 45, 0);
        vgpu_interaction& g = add_interaction(c, GENERAL, false, col(22));
        alu_tuple_scalar_out(g, weighted({{23, 104}, {24, 115}, {25, 117}, {26, 118}}), 0, 4, 21);
    }
    {   // 9 com32: output 11, is_ne 12, is_eq 13
        vgpu_chip_desc& c = def(9, 14, 0);
        vgpu_interaction& g = add_interaction(c, GENERAL, false, weighted({{12, 1}, {13, 1}}));
        alu_tuple_scalar_out(g, weighted({{12, 111}, {13, 116}}), 0, 4, 11);
    }
    {   // 10 bitwise32: out 72, is_and 76, is_or 77, is_xor 78
        vgpu_chip_desc& c = def(10, 79, 0);
        vgpu_interaction& g = add_interaction(c, GENERAL, false, weighted({{76, 1}, {77, 1}, {78, 1}}));
        alu_tuple(g, weighted({{76, 107}, {77, 108}, {78, 109}}), 0, 4, 72);
    }
    {   // 11 output: clk 0, value 1, is_real 2, opcode 6
        vgpu_chip_desc& c = def(11, 7, 0);
        vgpu_interaction& g = add_interaction(c, GENERAL, false, col(2));
        g.fields[g.n_fields++] = col(6);
        for (int i = 0; i < 12; i++) g.fields[g.n_fields++] = (i == 3) ? col(1) : konst(0);
        g.fields[g.n_fields++] = col(0);
    }
    {   // 12 range: mult 0, counter 1
        vgpu_chip_desc& c = def(12, 2, 1);
        vgpu_interaction& r = add_interaction(c, RANGE, false, col(0));
        r.fields[r.n_fields++] = col(1);
    }
    {   // 13 static_data: addr 0, value 1..4, is_real 5
        vgpu_chip_desc& c = def(13, 6, 0);
        vgpu_interaction& s = add_interaction(c, MEM, true, col(5));
        s.fields[s.n_fields++] = konst(0); s.fields[s.n_fields++] = konst(0); s.fields[s.n_fields++] = col(0); s.fields[s.n_fields++] = konst(1);
        word(s, 1);
    }
}

}  // namespace

extern "C" const vgpu_chip_desc* vgpu_basic_machine_chip(uint32_t chip_id) {
    std::call_once(g_once, init);
    return chip_id < VGPU_NUM_CHIPS ? &g_chips[chip_id] : nullptr;
}
