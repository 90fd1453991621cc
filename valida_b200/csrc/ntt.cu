// K1/K2 — batched BabyBear NTT / iNTT / coset-LDE over column-major device matrices (sm_100a).
// Replaces p3-dft's TwoAdicSubgroupDft::{dft_batch, idft_batch, coset_lde_batch} as reached from
// TwoAdicFriPcs::commit_shifted_batches (reference call sites derive/src/lib.rs:309,330,355,372;
// DFT selected at basic/src/bin/valida.rs:379).
//
// Design (B200-first, not the reference's row-major butterfly network):
//  * a length-n column transform is split n = n1*n2 ("four-step"); each pass stages a tile of
//    T sub-transforms of length L <= 2^12 in shared memory, runs all log2(L) radix-2 DIF stages
//    there, and applies the inter-pass twiddle / coset scaling on the way out, so a column crosses
//    HBM twice per transform regardless of n (<= 2^24 per coset);
//  * tiles are addressed as element(r, g) = base + r*rs + g*gs, with the thread->element map chosen
//    so that the unit-stride index is the fastest one in both the load and the store (T-wide
//    segments in the strided passes, whole rows in the contiguous pass);
//  * the coset LDE = iNTT (natural->natural, two tile passes with a transposed store) followed by
//    two forward coset transforms natural->bit-reversed (shift and shift*w_2n) written into the two
//    halves of the bit-reversed output — the committed row order needs no separate permutation;
//  * twiddles come from a two-level power table of w_(2^27) (lo[e&4095]*hi[e>>12]) kept L2-resident.
#include "ctx.h"

namespace {

using bb::mul; using bb::add; using bb::sub;

struct PassParams {
    const uint32_t* src; uint32_t* dst;
    uint64_t src_cs, dst_cs;          // column strides
    uint64_t src_rs, src_gs;          // element(r, g) = col + r*src_rs + g*src_gs
    uint64_t dst_rs, dst_gs;          // output k of group g -> col + pos*dst_rs + g*dst_gs
    uint32_t log_len;                 // L = 2^log_len (sub-transform length, staged in smem)
    uint32_t tile;                    // T groups per CTA
    uint64_t groups;                  // groups per column (n / L)
    uint32_t inverse;                 // twiddle direction
    uint32_t dst_natural;             // 1: output k stored at pos = k ; 0: raw DIF order pos = bitrev(k)
    // pre-multiplier (odd coset): x *= w_NMAX^((r*pre_r + g*pre_g) * pre_unit)
    uint32_t pre_mode; uint64_t pre_r, pre_g, pre_unit;
    // post-multiplier: mode 1: w_NMAX^(+-(g*k) * post_unit) ; mode 2: table(g*post_g + k*post_k) ; mode 3: constant scale
    uint32_t post_mode; uint64_t post_unit, post_g, post_k; uint32_t post_scale;
    const uint32_t* root_lo; const uint32_t* root_hi;   // w_NMAX tables
    const uint32_t* tab_lo; const uint32_t* tab_hi;     // shift tables (mode 2)
};

__device__ __forceinline__ uint32_t root_pow(const PassParams& p, uint64_t e) {
    e &= ((1ull << VG_LOG_NMAX) - 1);
    uint32_t lo = __ldg(p.root_lo + (e & (VG_POW_LO - 1)));
    uint32_t hi = __ldg(p.root_hi + (e >> VG_POW_LO_BITS));
    return mul(lo, hi);
}

// One CTA = one tile of `tile` sub-transforms of one column.  grid.x = tiles_per_col, grid.y = column.
__global__ void __launch_bounds__(1024) ntt_pass_kernel(PassParams p) {
    extern __shared__ uint32_t smem[];
    const uint32_t L = 1u << p.log_len, T = p.tile;
    const uint32_t LS = L + 1;                       // padded row pitch (bank spread for t-fastest access)
    uint32_t* tw = smem;                             // L/2 twiddles w_L^(+-j)
    uint32_t* data = smem + (L >= 2 ? L / 2 : 1);
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint64_t col = blockIdx.y;
    const uint64_t g0 = (uint64_t)blockIdx.x * T;
    const uint32_t* src = p.src + col * p.src_cs;
    uint32_t* dst = p.dst + col * p.dst_cs;

    // twiddles for the in-smem transform: w_L^j = w_NMAX^(j * NMAX/L); NMAX/L >= 2^15 so only the hi table is touched
    {
        const uint64_t unit = (1ull << VG_LOG_NMAX) >> p.log_len;
        for (uint32_t j = tid; j < L / 2; j += nt) {
            uint64_t e = (uint64_t)j * unit;
            if (p.inverse && e) e = (1ull << VG_LOG_NMAX) - e;
            tw[j] = __ldg(p.root_hi + (e >> VG_POW_LO_BITS));
        }
    }
    // ---- load ----
    const uint32_t total = L * T;
    if (p.src_gs == 1) {           // groups are adjacent in memory: t fastest
        for (uint32_t idx = tid; idx < total; idx += nt) {
            uint32_t t = idx % T, r = idx / T;
            uint32_t v = src[(uint64_t)r * p.src_rs + (g0 + t)];
            if (p.pre_mode) v = mul(v, root_pow(p, ((uint64_t)r * p.pre_r + (g0 + t) * p.pre_g) * p.pre_unit));
            data[t * LS + r] = v;
        }
    } else {                       // each group is contiguous: r fastest
        for (uint32_t idx = tid; idx < total; idx += nt) {
            uint32_t r = idx & (L - 1), t = idx >> p.log_len;
            uint32_t v = src[(uint64_t)r * p.src_rs + (g0 + t) * p.src_gs];
            if (p.pre_mode) v = mul(v, root_pow(p, ((uint64_t)r * p.pre_r + (g0 + t) * p.pre_g) * p.pre_unit));
            data[t * LS + r] = v;
        }
    }
    __syncthreads();
    // ---- log2(L) DIF stages, natural in -> bit-reversed positions out ----
    const uint32_t nbf = total >> 1;   // butterflies per stage
    for (uint32_t s = 0; s < p.log_len; s++) {
        const uint32_t lh = p.log_len - 1 - s;      // log2(half)
        const uint32_t half = 1u << lh;
        for (uint32_t b = tid; b < nbf; b += nt) {
            uint32_t t = b >> (p.log_len - 1), bi = b & ((L >> 1) - 1);
            uint32_t j = bi & (half - 1), blk = bi >> lh;
            uint32_t i0 = t * LS + (blk << (lh + 1)) + j, i1 = i0 + half;
            uint32_t a = data[i0], c = data[i1];
            data[i0] = add(a, c);
            uint32_t d = sub(a, c);
            data[i1] = lh == 0 ? d : mul(d, tw[j << s]);   // j == 0 in the last stage: twiddle 1
        }
        __syncthreads();
    }
    // ---- store (with optional post multiplier) ----
    auto post = [&](uint32_t v, uint32_t k, uint64_t g) -> uint32_t {
        if (p.post_mode == 1) {
            uint64_t e = (g * k) * p.post_unit;
            if (p.inverse && e) e = (1ull << VG_LOG_NMAX) - (e & ((1ull << VG_LOG_NMAX) - 1));
            return mul(v, root_pow(p, e));
        } else if (p.post_mode == 2) {
            uint64_t e = g * p.post_g + (uint64_t)k * p.post_k;
            return mul(v, mul(__ldg(p.tab_lo + (e & (VG_POW_LO - 1))), __ldg(p.tab_hi + (e >> VG_POW_LO_BITS))));
        } else if (p.post_mode == 3) {
            return mul(v, p.post_scale);
        }
        return v;
    };
    if (p.dst_gs == 1) {           // t fastest
        for (uint32_t idx = tid; idx < total; idx += nt) {
            uint32_t t = idx % T, pos = idx / T;
            uint32_t q = p.dst_natural ? bb::reverse_bits(pos, p.log_len) : pos;   // smem slot holding the value for `pos`
            uint32_t k = p.dst_natural ? pos : bb::reverse_bits(pos, p.log_len);   // its natural output index
            dst[(uint64_t)pos * p.dst_rs + (g0 + t)] = post(data[t * LS + q], k, g0 + t);
        }
    } else {                       // pos fastest
        for (uint32_t idx = tid; idx < total; idx += nt) {
            uint32_t pos = idx & (L - 1), t = idx >> p.log_len;
            uint32_t q = p.dst_natural ? bb::reverse_bits(pos, p.log_len) : pos;
            uint32_t k = p.dst_natural ? pos : bb::reverse_bits(pos, p.log_len);
            dst[(uint64_t)pos * p.dst_rs + (g0 + t) * p.dst_gs] = post(data[t * LS + q], k, g0 + t);
        }
    }
}

constexpr int LOG_LMAX = 12;   // largest sub-transform staged in shared memory

struct Split { int l1, l2; };
Split choose_split(int log_n) {
    if (log_n <= LOG_LMAX) return {log_n, 0};
    int l1 = (log_n + 1) / 2;
    return {l1, log_n - l1};
}
uint32_t choose_tile(int log_len, uint64_t groups) {
    // L*T <= 2^14 elements (64 KB) so that 2-3 CTAs co-reside per SM; T <= 32 (128 B segments)
    uint32_t t = 1u << (14 - log_len > 5 ? 5 : (14 - log_len < 0 ? 0 : 14 - log_len));
    while (t > groups) t >>= 1;
    return t ? t : 1;
}

int32_t launch_pass(vgpu_ctx* ctx, PassParams p, uint64_t w) {
    p.root_lo = ctx->root_table.lo; p.root_hi = ctx->root_table.hi;
    uint32_t L = 1u << p.log_len;
    p.tile = choose_tile((int)p.log_len, p.groups);
    uint64_t tiles = p.groups / p.tile;
    size_t smem = ((L >= 2 ? L / 2 : 1) + (size_t)p.tile * (L + 1)) * sizeof(uint32_t);
    uint32_t total = L * p.tile;
    uint32_t threads = total / 2 >= 1024 ? 1024 : (total / 2 >= 32 ? total / 2 : 32);
    if (threads > 512 && smem <= 70 * 1024) threads = 512;
    static bool attr_set = false;
    if (!attr_set) { VG_CUDA(ctx, cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr_set = true; }
    // grid.y limit 65535: chunk columns
    for (uint64_t c0 = 0; c0 < w; c0 += 65535) {
        uint64_t wc = w - c0 < 65535 ? w - c0 : 65535;
        PassParams q = p;
        q.src = p.src + c0 * p.src_cs; q.dst = p.dst + c0 * p.dst_cs;
        dim3 grid((unsigned)tiles, (unsigned)wc);
        KScope ks(ctx, KC_NTT, 8.0 * (double)(p.groups << p.log_len) * (double)wc);
        ntt_pass_kernel<<<grid, threads, smem, ctx->stream>>>(q);
        VG_LAUNCH_CHECK(ctx);
    }
    return 0;
}

__global__ void zero_pad_kernel(const uint32_t* src, uint64_t src_cs, uint32_t* dst, uint64_t dst_cs, uint64_t h, uint64_t H, uint64_t w) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * w) return;
    uint64_t c = i / H, r = i % H;
    dst[c * dst_cs + r] = r < h ? src[c * src_cs + r] : 0;
}

}  // namespace

// natural -> natural transform of every column (forward: out[k] = sum_j in[j] w^(jk); inverse scales by 1/n).
// coset != null (inverse only): coefficient k additionally multiplied by shift^k (table carries the 1/n).
int32_t vg_ntt_nat2nat(vgpu_ctx* ctx, const uint32_t* src, uint64_t src_cs, uint32_t* dst, uint64_t dst_cs, int log_n, uint64_t w,
                       bool inverse, const PowTable* coset, uint32_t* tmp, uint64_t tmp_cs) {
    const uint64_t n = 1ull << log_n;
    Split sp = choose_split(log_n);
    uint32_t ninv = inverse ? bb::inv(bb::to_monty((uint32_t)(n % bb::P))) : bb::R1;
    PassParams p{};
    p.inverse = inverse;
    if (sp.l2 == 0) {
        p.src = src; p.src_cs = src_cs; p.dst = dst; p.dst_cs = dst_cs;
        p.src_rs = 1; p.src_gs = n; p.dst_rs = 1; p.dst_gs = n;
        p.log_len = log_n; p.groups = 1; p.dst_natural = 1;
        if (coset) { p.post_mode = 2; p.post_g = 0; p.post_k = 1; p.tab_lo = coset->lo; p.tab_hi = coset->hi; }
        else if (inverse) { p.post_mode = 3; p.post_scale = ninv; }
        return launch_pass(ctx, p, w);
    }
    const uint64_t n1 = 1ull << sp.l1, n2 = 1ull << sp.l2;
    // pass 1: over i1 (stride n2) for each i2; times w_n^(+-i2*k1); transposed store tmp[i2][k1]
    p.src = src; p.src_cs = src_cs; p.dst = tmp; p.dst_cs = tmp_cs;
    p.src_rs = n2; p.src_gs = 1; p.dst_rs = 1; p.dst_gs = n1;
    p.log_len = sp.l1; p.groups = n2; p.dst_natural = 1;
    p.post_mode = 1; p.post_unit = (1ull << VG_LOG_NMAX) >> log_n;
    VG_TRY(launch_pass(ctx, p, w));
    // pass 2: tmp is [i2][k1]; over i2 (stride n1) for each k1; X[k1 + n1*k2] stored at that index
    PassParams q{};
    q.inverse = inverse;
    q.src = tmp; q.src_cs = tmp_cs; q.dst = dst; q.dst_cs = dst_cs;
    q.src_rs = n1; q.src_gs = 1; q.dst_rs = n1; q.dst_gs = 1;
    q.log_len = sp.l2; q.groups = n1; q.dst_natural = 1;
    if (coset) { q.post_mode = 2; q.post_g = 1; q.post_k = n1; q.tab_lo = coset->lo; q.tab_hi = coset->hi; }
    else if (inverse) { q.post_mode = 3; q.post_scale = ninv; }
    return launch_pass(ctx, q, w);
}

// forward coset transform, natural coefficients -> bit-reversed evaluations; odd != 0 multiplies
// coefficient i by w_2n^i first (the second coset of the blow-up-2 LDE).
static int32_t ntt_nat2bitrev(vgpu_ctx* ctx, const uint32_t* src, uint64_t src_cs, uint32_t* dst, uint64_t dst_cs, int log_n, uint64_t w, bool odd) {
    const uint64_t n = 1ull << log_n;
    Split sp = choose_split(log_n);
    PassParams p{};
    p.inverse = 0;
    if (odd) { p.pre_mode = 1; p.pre_unit = (1ull << VG_LOG_NMAX) >> (log_n + 1); }
    if (sp.l2 == 0) {
        p.src = src; p.src_cs = src_cs; p.dst = dst; p.dst_cs = dst_cs;
        p.src_rs = 1; p.src_gs = n; p.dst_rs = 1; p.dst_gs = n;
        p.log_len = log_n; p.groups = 1; p.dst_natural = 0;
        p.pre_r = 1; p.pre_g = 0;
        return launch_pass(ctx, p, w);
    }
    const uint64_t n1 = 1ull << sp.l1, n2 = 1ull << sp.l2;
    p.src = src; p.src_cs = src_cs; p.dst = dst; p.dst_cs = dst_cs;
    p.src_rs = n2; p.src_gs = 1; p.dst_rs = n2; p.dst_gs = 1;
    p.log_len = sp.l1; p.groups = n2; p.dst_natural = 0;
    p.pre_r = n2; p.pre_g = 1;
    p.post_mode = 1; p.post_unit = (1ull << VG_LOG_NMAX) >> log_n;
    VG_TRY(launch_pass(ctx, p, w));
    PassParams q{};
    q.src = dst; q.src_cs = dst_cs; q.dst = dst; q.dst_cs = dst_cs;
    q.src_rs = 1; q.src_gs = n2; q.dst_rs = 1; q.dst_gs = n2;
    q.log_len = sp.l2; q.groups = n1; q.dst_natural = 0;
    return launch_pass(ctx, q, w);
}

// coset_lde_batch(mat, added_bits = 1, shift): dst (2h rows per column).
int32_t vg_coset_lde(vgpu_ctx* ctx, const uint32_t* src, uint64_t src_cs, uint64_t h, uint64_t w, uint32_t shift_canonical,
                     uint32_t* dst, uint64_t dst_cs, bool bit_reversed) {
    int log_n = 0;
    while ((1ull << log_n) < h) log_n++;
    if ((1ull << log_n) != h) VG_FAIL(ctx, "coset_lde: height %llu is not a power of two", (unsigned long long)h);
    if (log_n + 1 > VG_LOG_NMAX) VG_FAIL(ctx, "coset_lde: LDE height 2^%d exceeds BabyBear two-adicity", log_n + 1);
    const PowTable* tab = nullptr;
    uint32_t ninv_canon = bb::from_monty(bb::inv(bb::to_monty((uint32_t)(h % bb::P))));
    VG_TRY(vg_get_shift_table(ctx, shift_canonical, ninv_canon, h, &tab));
    // column batches: bound scratch (coefficients + transposed intermediate) to ~512 MB and keep
    // mid-size batches L2-resident between passes
    uint64_t batch = (uint64_t)(24u << 20) / (4 * h);
    if (batch < 1) batch = 1;
    if (batch > w) batch = w;
    uint32_t *coef = nullptr, *tmp = nullptr;
    VG_TRY(vg_alloc(ctx, (void**)&coef, batch * h * 4));
    VG_TRY(vg_alloc(ctx, (void**)&tmp, batch * h * 4));
    int32_t rc = 0;
    for (uint64_t c0 = 0; c0 < w && rc == 0; c0 += batch) {
        uint64_t wc = w - c0 < batch ? w - c0 : batch;
        rc = vg_ntt_nat2nat(ctx, src + c0 * src_cs, src_cs, coef, h, log_n, wc, true, tab, tmp, h);
        if (rc) break;
        if (bit_reversed) {
            rc = ntt_nat2bitrev(ctx, coef, h, dst + c0 * dst_cs, dst_cs, log_n, wc, false);
            if (rc) break;
            rc = ntt_nat2bitrev(ctx, coef, h, dst + c0 * dst_cs + h, dst_cs, log_n, wc, true);
        } else {
            // natural-order output (API completeness, not on the proving path): zero-pad and transform at size 2h
            uint32_t* pad = nullptr; uint32_t* tmp2 = nullptr;
            rc = vg_alloc(ctx, (void**)&pad, wc * 2 * h * 4); if (rc) break;
            rc = vg_alloc(ctx, (void**)&tmp2, wc * 2 * h * 4); if (rc) { vg_free(ctx, pad); break; }
            uint64_t tot = 2 * h * wc;
            zero_pad_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(coef, h, pad, 2 * h, h, 2 * h, wc);
            ctx->launches++;
            rc = vg_ntt_nat2nat(ctx, pad, 2 * h, dst + c0 * dst_cs, dst_cs, log_n + 1, wc, false, nullptr, tmp2, 2 * h);
            vg_free(ctx, pad); vg_free(ctx, tmp2);
        }
    }
    vg_free(ctx, coef); vg_free(ctx, tmp);
    return rc;
}
