"""Synthetic multi-chip programs shared by the CPU and GPU tests (a loop mixing add, sub,
lt/lte/slt/sle, and/or/xor on a linear-congruential value stream; mul/div/shift are excluded there because the
reference's range-bus accounting does not balance for them, and eq/ne because the reference's Com32Chip::op_to_row
fills only the opcode flags — alu_u32/src/com/mod.rs:84-100 — so its bus receive can never match the CPU's send)."""
import numpy as np

B = 24


def mixed_program(iters):
    """cpu + mem + add + lt + range + program: add / lt / lte / slt / sle (incl. left immediates), bne back-edge."""
    return np.array([
        [7, -4, 0, 0, 0, 0],                                  # i = 0
        [7, -8, 0x12, 0x34, 0x56, 0x78],                      # x = seed
        [100, -8, -8, 1013904223, 0, 1],                      # x += c            (add32 imm)
        [100, -12, -8, -4, 0, 0],                             # y = x + i         (add32)
        [104, -16, -12, -8, 0, 0],                            # y < x             (lt32)
        [117, -20, -8, -12, 0, 0],                            # x <s y            (slt32)
        [115, -24, 77, -8, 1, 0],                             # 77 <= x           (lte32, left immediate)
        [118, -28, -12, 1000, 0, 1],                          # y <=s 1000        (sle32, right immediate)
        [100, -4, -4, 1, 0, 1],                               # i += 1
        [6, 2 * B, -4, iters, 0, 1],                          # bne loop, i, iters
        [8, 0, 0, 0, 0, 0],
    ], dtype=np.int32)


def config5_program(iters):
    """cpu + mem + add + sub + lt + bitwise + range + program.  x <- x*1664525 + 1013904223 is replaced by an
    add/xor/and/or mix (no mul); every ALU flavour appears with a memory operand and with an immediate.
    The sub operands are shaped so that no byte borrows (minuend bytes >= 0x80 > subtrahend bytes): the reference's
    Sub32Chip::op_to_row derives each borrow from the operand bytes alone (alu_u32/src/sub/mod.rs:103-111) and its AIR
    has no top-byte borrow (sub/stark.rs:42-45), so a subtraction that borrows is unprovable in the reference too."""
    return np.array([
        [7, -4, 0, 0, 0, 0],                                  # i = 0
        [7, -8, 0x9e, 0x37, 0x79, 0xb9],                      # x = seed
        [7, -32, 0x0f, 0xf0, 0x55, 0xaa],                     # m = mask
        [100, -8, -8, 1013904223, 0, 1],                      # x += c                       (add32 imm)     <- loop
        [109, -12, -8, -32, 0, 0],                            # y = x ^ m                    (xor32)
        [107, -16, -12, 0x7f7f7f7f, 0, 1],                    # t = y & 0x7f7f7f7f           (and32 imm)
        [108, -20, -8, -0x7f7f7f80, 0, 1],                    # s = x | 0x80808080           (or32 imm)
        [101, -24, -20, -16, 0, 0],                           # d = s - t                    (sub32)
        [101, -28, -20, 12345, 0, 1],                         # e = s - 12345                (sub32 imm)
        [109, -8, -8, -28, 0, 0],                             # x ^= e                       (xor32)
        [104, -36, -24, -8, 0, 0],                            # d < x                        (lt32)
        [118, -40, -28, -12, 0, 0],                           # e <=s y                      (sle32)
        [117, -44, 5, -28, 1, 0],                             # 5 <s e                       (slt32 left imm)
        [115, -48, -20, 4096, 0, 1],                          # s <= 4096                    (lte32 imm)
        [107, -32, -32, -12, 0, 0],                           # m &= y                       (and32)
        [108, -32, -32, 0x01010101, 0, 1],                    # m |= 0x01010101              (or32 imm)
        [100, -4, -4, 1, 0, 1],                               # i += 1
        [6, 3 * B, -4, iters, 0, 1],                          # bne loop, i, iters
        [8, 0, 0, 0, 0, 0],
    ], dtype=np.int32)


def static_data_program():
    """prove_static_data of the reference (basic/tests/test_static_data.rs:30-55): loops forever unless the static value is loaded.
        _start: imm32 0(fp), 0, 0, 0, 0x10 ; load32 -4(fp), 0(fp) ; bnei _start, -4(fp), 0x25 ; stop
    with static cells 0x10 = Word([0,0,0,0x25]), 0x14 = Word([0,0,0,0x32]) (test_static_data.rs:59-60)."""
    prog = np.array([
        [7, 0, 0, 0, 0, 0x10],
        [1, -4, 0, 0, 0, 0],
        [6, 0, -4, 0x25, 0, 1],
        [8, 0, 0, 0, 0, 0],
    ], dtype=np.int32)
    return prog, {0x10: 0x25, 0x14: 0x32}
