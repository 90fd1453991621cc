// Keccak-f[1600] on 32-bit register pairs for sm_100a (LOP3 for theta/chi, SHF funnel shifts for rho).
// Replaces tiny-keccak's keccakf as used by p3-keccak::Keccak256Hash inside
// SerializingHasher32 / CompressionFunctionFromHasher (basic/src/bin/valida.rs:367-371).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace kk {

__constant__ uint2 RC[24] = {
    {0x00000001u, 0x00000000u}, {0x00008082u, 0x00000000u}, {0x0000808au, 0x80000000u}, {0x80008000u, 0x80000000u},
    {0x0000808bu, 0x00000000u}, {0x80000001u, 0x00000000u}, {0x80008081u, 0x80000000u}, {0x00008009u, 0x80000000u},
    {0x0000008au, 0x00000000u}, {0x00000088u, 0x00000000u}, {0x80008009u, 0x00000000u}, {0x8000000au, 0x00000000u},
    {0x8000808bu, 0x00000000u}, {0x0000008bu, 0x80000000u}, {0x00008089u, 0x80000000u}, {0x00008003u, 0x80000000u},
    {0x00008002u, 0x80000000u}, {0x00000080u, 0x80000000u}, {0x0000800au, 0x00000000u}, {0x8000000au, 0x80000000u},
    {0x80008081u, 0x80000000u}, {0x00008080u, 0x80000000u}, {0x80000001u, 0x00000000u}, {0x80008008u, 0x80000000u}};

// x = lo, y = hi
template <int N> __device__ __forceinline__ uint2 rol(uint2 a) {
    if (N == 0) return a;
    if (N == 32) return make_uint2(a.y, a.x);
    if (N < 32) return make_uint2(__funnelshift_l(a.y, a.x, N), __funnelshift_l(a.x, a.y, N));
    return make_uint2(__funnelshift_l(a.x, a.y, N - 32), __funnelshift_l(a.y, a.x, N - 32));
}
__device__ __forceinline__ uint2 x2(uint2 a, uint2 b) { return make_uint2(a.x ^ b.x, a.y ^ b.y); }
__device__ __forceinline__ uint2 x5(uint2 a, uint2 b, uint2 c, uint2 d, uint2 e) { return make_uint2(a.x ^ b.x ^ c.x ^ d.x ^ e.x, a.y ^ b.y ^ c.y ^ d.y ^ e.y); }
__device__ __forceinline__ uint2 chi(uint2 a, uint2 b, uint2 c) { return make_uint2(a.x ^ (~b.x & c.x), a.y ^ (~b.y & c.y)); }

// State lane (x, y) lives in A[x + 5 y].
__device__ __forceinline__ void keccak_f(uint2 A[25]) {
#pragma unroll 1
    for (int round = 0; round < 24; round++) {
        uint2 C0 = x5(A[0], A[5], A[10], A[15], A[20]);
        uint2 C1 = x5(A[1], A[6], A[11], A[16], A[21]);
        uint2 C2 = x5(A[2], A[7], A[12], A[17], A[22]);
        uint2 C3 = x5(A[3], A[8], A[13], A[18], A[23]);
        uint2 C4 = x5(A[4], A[9], A[14], A[19], A[24]);
        uint2 D0 = x2(C4, rol<1>(C1)), D1 = x2(C0, rol<1>(C2)), D2 = x2(C1, rol<1>(C3)), D3 = x2(C2, rol<1>(C4)), D4 = x2(C3, rol<1>(C0));
        // theta + rho + pi:  B[y, 2x+3y] = rol(A[x,y] ^ D[x], r[x,y])
        uint2 B0 = x2(A[0], D0);
        uint2 B10 = rol<1>(x2(A[1], D1));
        uint2 B20 = rol<62>(x2(A[2], D2));
        uint2 B5 = rol<28>(x2(A[3], D3));
        uint2 B15 = rol<27>(x2(A[4], D4));
        uint2 B16 = rol<36>(x2(A[5], D0));
        uint2 B1 = rol<44>(x2(A[6], D1));
        uint2 B11 = rol<6>(x2(A[7], D2));
        uint2 B21 = rol<55>(x2(A[8], D3));
        uint2 B6 = rol<20>(x2(A[9], D4));
        uint2 B7 = rol<3>(x2(A[10], D0));
        uint2 B17 = rol<10>(x2(A[11], D1));
        uint2 B2 = rol<43>(x2(A[12], D2));
        uint2 B12 = rol<25>(x2(A[13], D3));
        uint2 B22 = rol<39>(x2(A[14], D4));
        uint2 B23 = rol<41>(x2(A[15], D0));
        uint2 B8 = rol<45>(x2(A[16], D1));
        uint2 B18 = rol<15>(x2(A[17], D2));
        uint2 B3 = rol<21>(x2(A[18], D3));
        uint2 B13 = rol<8>(x2(A[19], D4));
        uint2 B14 = rol<18>(x2(A[20], D0));
        uint2 B24 = rol<2>(x2(A[21], D1));
        uint2 B9 = rol<61>(x2(A[22], D2));
        uint2 B19 = rol<56>(x2(A[23], D3));
        uint2 B4 = rol<14>(x2(A[24], D4));
        // chi
        A[0] = chi(B0, B1, B2); A[1] = chi(B1, B2, B3); A[2] = chi(B2, B3, B4); A[3] = chi(B3, B4, B0); A[4] = chi(B4, B0, B1);
        A[5] = chi(B5, B6, B7); A[6] = chi(B6, B7, B8); A[7] = chi(B7, B8, B9); A[8] = chi(B8, B9, B5); A[9] = chi(B9, B5, B6);
        A[10] = chi(B10, B11, B12); A[11] = chi(B11, B12, B13); A[12] = chi(B12, B13, B14); A[13] = chi(B13, B14, B10); A[14] = chi(B14, B10, B11);
        A[15] = chi(B15, B16, B17); A[16] = chi(B16, B17, B18); A[17] = chi(B17, B18, B19); A[18] = chi(B18, B19, B15); A[19] = chi(B19, B15, B16);
        A[20] = chi(B20, B21, B22); A[21] = chi(B21, B22, B23); A[22] = chi(B22, B23, B24); A[23] = chi(B23, B24, B20); A[24] = chi(B24, B20, B21);
        // iota
        uint2 rc = RC[round];
        A[0].x ^= rc.x; A[0].y ^= rc.y;
    }
}

}  // namespace kk
