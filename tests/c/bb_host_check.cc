// Host instantiation of the product's field arithmetic (valida_b200/csrc/bb.cuh — the text the kernels compile, minus the
// __CUDA_ARCH__ branches) driven from stdin: one operation per line, operands as decimal Montgomery words, result on stdout.
// Built with g++ by tests/test_bb_host_arith.py, which checks every answer against Python integers.  No GPU, no CUDA runtime call.
#include <cstdio>
#include <cstring>
#include "bb.cuh"

int main() {
    char op[32];
    while (scanf("%31s", op) == 1) {
        if (!strcmp(op, "mul") || !strcmp(op, "add") || !strcmp(op, "sub")) {
            unsigned a, b; if (scanf("%u %u", &a, &b) != 2) return 1;
            unsigned r = !strcmp(op, "mul") ? bb::mul(a, b) : !strcmp(op, "add") ? bb::add(a, b) : bb::sub(a, b);
            printf("%u\n", r);
        } else if (!strcmp(op, "neg") || !strcmp(op, "inv") || !strcmp(op, "to_monty") || !strcmp(op, "from_monty")) {
            unsigned a; if (scanf("%u", &a) != 1) return 1;
            unsigned r = !strcmp(op, "neg") ? bb::neg(a) : !strcmp(op, "inv") ? bb::inv(a) : !strcmp(op, "to_monty") ? bb::to_monty(a) : bb::from_monty(a);
            printf("%u\n", r);
        } else if (!strcmp(op, "reduce64")) {
            unsigned long long t; if (scanf("%llu", &t) != 1) return 1;
            printf("%u\n", bb::monty_reduce64(t));
        } else if (!strcmp(op, "pow")) {
            unsigned a; unsigned long long e; if (scanf("%u %llu", &a, &e) != 2) return 1;
            printf("%u\n", bb::pow(a, e));
        } else if (!strcmp(op, "revbits")) {
            unsigned x; int bits; if (scanf("%u %d", &x, &bits) != 2) return 1;
            printf("%u\n", bb::reverse_bits(x, bits));
        } else if (!strcmp(op, "gen")) {
            int bits; if (scanf("%d", &bits) != 1) return 1;
            printf("%u\n", bb::two_adic_generator_monty(bits));
        } else if (!strcmp(op, "e5mul") || !strcmp(op, "e5add") || !strcmp(op, "e5sub")) {
            bb::E5 a, b;
            for (int i = 0; i < 5; i++) if (scanf("%u", &a.c[i]) != 1) return 1;
            for (int i = 0; i < 5; i++) if (scanf("%u", &b.c[i]) != 1) return 1;
            bb::E5 r = !strcmp(op, "e5mul") ? bb::e5_mul(a, b) : !strcmp(op, "e5add") ? bb::e5_add(a, b) : bb::e5_sub(a, b);
            printf("%u %u %u %u %u\n", r.c[0], r.c[1], r.c[2], r.c[3], r.c[4]);
        } else if (!strcmp(op, "e5inv") || !strcmp(op, "e5frob")) {
            bb::E5 a;
            for (int i = 0; i < 5; i++) if (scanf("%u", &a.c[i]) != 1) return 1;
            bb::E5 r;
            if (!strcmp(op, "e5inv")) r = bb::e5_inv(a);
            else { uint32_t z[5]; bb::e5_frob_consts(z); r = bb::e5_frobenius(a, z); }
            printf("%u %u %u %u %u\n", r.c[0], r.c[1], r.c[2], r.c[3], r.c[4]);
        } else return 2;
    }
    return 0;
}
