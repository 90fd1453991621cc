/* A caller of the C ABI with no Python in the process — the stand-in, in this toolchain, for the Rust FFI crate of
 * runs fib_program(n) of basic/tests/test_prover.rs:35-188 through the host witness generator, proves it with
 * vgpu_prove (host buffers, exactly what `RowMajorMatrix<Val>.values` pointers would be), checks the proof with
 * vgpu_verify and writes the CBOR bytes to argv[2].
 *   gcc -O2 -I include tests/c/c_abi_smoke.c -o c_abi_smoke -L valida_b200 -lvalida_b200 -Wl,-rpath,$PWD/valida_b200
 *   ./c_abi_smoke 25 proof.cbor
 * Exit codes: 0 proved and verified; 3 no CUDA device (the library has no CPU fallback); 1 anything else. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "valida_b200.h"

/* the documented stand-in for the caller's Poseidon RNG: SplitMix64 seeded with ASCII "valida", 31-bit rejection sampling */
static void round_constants(uint32_t rc[480]) {
    uint64_t state = 0x76616C696461ull;
    int k = 0;
    while (k < 480) {
        state += 0x9E3779B97F4A7C15ull;
        uint64_t z = state;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        uint64_t c = z >> 33;
        if (c < 2013265921ull) rc[k++] = (uint32_t)c;
    }
}

int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? (uint32_t)strtoul(argv[1], NULL, 10) : 25;
    int32_t program[23 * 6];
    const uint64_t n_instr = vgpu_fib_program(n, program);
    char err[256] = {0};
    vgpu_traces* t = NULL;
    if (vgpu_machine_run(program, n_instr, 0, 0x1000, 1ull << 30, &t, err, sizeof err) != 0) { fprintf(stderr, "machine_run: %s\n", err); return 1; }
    uint32_t clock = 0, mem_ops = 0, add_ops = 0;
    vgpu_traces_stats(t, &clock, &mem_ops, &add_ops);
    printf("fib(%u): %u cycles, %u memory operations, %u additions\n", n, clock, mem_ops, add_ops);

    vgpu_ctx* ctx = NULL;
    int32_t rc = vgpu_ctx_create(0, NULL, &ctx);
    if (rc != 0) {
        fprintf(stderr, "ctx_create: %s\n", ctx ? vgpu_last_error(ctx) : "allocation failed");
        if (ctx) vgpu_ctx_destroy(ctx);
        vgpu_traces_free(t);
        return rc == -2 ? 3 : 1;
    }
    uint32_t rcs[480];
    round_constants(rcs);
    if (vgpu_set_challenger(ctx, rcs, NULL) != 0) { fprintf(stderr, "set_challenger: %s\n", vgpu_last_error(ctx)); return 1; }
    vgpu_matrix main_traces[VGPU_NUM_CHIPS], prep[2];
    for (uint32_t i = 0; i < VGPU_NUM_CHIPS; i++) main_traces[i] = *vgpu_traces_main(t, i);
    for (uint32_t i = 0; i < 2; i++) prep[i] = *vgpu_traces_preprocessed(t, i);
    uint8_t* proof = NULL;
    uint64_t len = 0;
    if (vgpu_prove(ctx, main_traces, prep, VGPU_REPR_CANONICAL, &proof, &len) != 0) { fprintf(stderr, "prove: %s\n", vgpu_last_error(ctx)); return 1; }
    int32_t verdict = -1;
    if (vgpu_verify(ctx, proof, len, prep, VGPU_REPR_CANONICAL, &verdict) != 0) { fprintf(stderr, "verify: %s\n", vgpu_last_error(ctx)); return 1; }
    printf("proof: %llu bytes, verdict %d, %llu kernel launches\n", (unsigned long long)len, verdict, (unsigned long long)vgpu_ctx_launch_count(ctx));
    if (argc > 2) {
        FILE* f = fopen(argv[2], "wb");
        if (!f || fwrite(proof, 1, len, f) != len) { fprintf(stderr, "cannot write %s\n", argv[2]); return 1; }
        fclose(f);
    }
    vgpu_free_bytes(proof);
    vgpu_traces_free(t);
    vgpu_ctx_destroy(ctx);
    return verdict == VGPU_ACCEPT ? 0 : 1;
}
