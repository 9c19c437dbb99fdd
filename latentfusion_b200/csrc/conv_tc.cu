// tcgen05 implicit-GEMM convolution for sm_100a — the dense contraction of the LatentFusion hot path
// (3x3x3 / 3x3 / 1x1 Equalized convs, channels-last) on the 5th-gen tensor cores with the
// He-scale + bias + LeakyReLU + PixelNorm epilogue fused.
//
// GEMM view per CTA step:  D[128 positions, Cout] += A_tap[128 positions, Cin] * W_tap[Cin, Cout]
//   * A is never im2col'ed.  A strip of (R + 2h) input rows x (W + 2h) columns of one depth plane is staged
//     once in shared memory as bf16 in the UMMA *no-swizzle K-major* canonical layout
//         [k-chunk of 8 channels][flattened padded position][8 x bf16]          (one 16-byte row per position)
//     With that layout a filter tap (dz,dy,dx) is nothing but a different descriptor START ADDRESS
//     (+ (dy*P + dx) * 16 bytes inside the plane, another ring slot for dz): the 27 taps of a 3x3x3 filter
//     are 27 tcgen05.mma instructions reading the same staged bytes.  M-tiles are 128 consecutive flattened
//     positions of the padded-pitch space; the <= 2h garbage columns per row are dropped in the epilogue.
//   * The CTA marches along depth with a ring of staged planes, so every input plane is read from
//     L2/HBM once per strip (plus the row halo), converted fp32 -> bf16 on the fly by 4 producer warps.
//   * Accumulators live in TMEM (double buffered, NT tiles x Cout columns each); one elected thread issues
//     the MMAs; 4 epilogue warps read TMEM with tcgen05.ld (one output position = one thread = all Cout
//     channels in registers, so PixelNorm is a thread-local reduction) and write channels-last fp32.
//   * Precision: operands are bf16, accumulation fp32.  "bf16x3" (precision 1) is hi*hi + lo*hi + hi*lo
//     (x = hi + lo, both bf16) => ~2^-16 relative per product, i.e. fp32-parity grade; "bf16" (precision 2) is the
//     single hi*hi product.  2-D layers do all three products in ONE launch (Params::tri: hi and lo slabs side by
//     side in every ring slot, x_hi * [W_hi | W_lo] as one N = 2*Cout MMA and x_lo * W_hi as an N = Cout MMA on the
//     same weight tile, both accumulator halves summed in the epilogue); 3-D layers that reach this kernel and
//     layers too large for that run two passes (dual weights) or three, accumulating in the fp32 output.
//   * The packed weights arrive by one bulk-TMA copy (cp.async.bulk + mbarrier) while the producers stage the first slab.
//
// Pipelines (mbarriers): slab_full/slab_empty[ring] (producers <-> MMA), acc_full/acc_empty[2] (MMA <-> epilogue).
// Every wait is bounded: a pipeline bug traps instead of hanging the GPU.
#include "common.cuh"

#include <cuda_bf16.h>
#include <stdlib.h>

namespace lf {
namespace tc {

constexpr int kMmaWarps = 5;                               // at most; issuer warp m owns M-tiles m, m+nmma, ...
constexpr int kEpiWarps = 4;                               // one warp per TMEM lane quarter (8 measured slower)
constexpr int kWarps = 16;
constexpr int kThreads = 32 * kWarps;                      // producers | MMA issuers | epilogue (last 4 warps)
// Roles are assigned at run time: nmma = min(NT, 5) issuer warps (one per M-tile), 4 epilogue warps, and every
// remaining warp (7..11) is a producer — a plan with few tiles gets more loads in flight instead of idle issuers.
constexpr int kMaxRing = 4;
constexpr int kMaxTiles = 5;

enum PassMode { PASS_ONLY = 0, PASS_FIRST = 1, PASS_MID = 2, PASS_LAST = 3 };

struct Params {
    const float* x; const uint16_t* wpk; const float* bias; float* y; float* rnorm;
    // fused backward prologue (bwd-data of a Block conv): x = upstream grad g, pro_y = the layer's output y,
    // pro_r = its saved PixelNorm denominators; the producers stage du = LeakyReLU'(y) * PixelNorm^T(g) on the fly
    const float* pro_y; const float* pro_r; int pro_act, pro_norm, pro_lg; float pro_slope;
    // fused backward EPILOGUE (bwd-data feeding a Block conv): the result row gx is pushed through the
    // PixelNorm/LeakyReLU backward of the layer that produced this conv's input (epi_y = that layer's output,
    // epi_r = its saved norms), i.e. the kernel writes du_prev instead of gx and the separate pass disappears
    const float* epi_y; const float* epi_r; int epi_act, epi_norm; float epi_slope;
    int n, d, h, w;            // extent (d = 1 for 2-D)
    int cin, cout, cin_pad, cout_pad;
    int k, hz;                 // kernel size (1|3); hz = depth halo (k/2 for 3-D, 0 for 2-D)
    int R, NT, P, pos_alloc, ring, DC;
    int nstrips, ndchunks, items;
    float scale; int act; float slope; int norm;
    int a_part;                // 0: hi = bf16(x), 1: lo = bf16(x - hi)
    int pass_mode;
    int nprod, nmma;           // producer / MMA-issuer warp counts (nprod + nmma + 4 == 16)
    int dual, ncols;           // dual: B = [W_hi | W_lo] (N = 2*cout_pad): one pass yields x_hi*W_hi + x_hi*W_lo; ncols = MMA N
    // tri (with dual): the whole bf16x3 product in ONE pass.  Every ring slot holds the hi slab followed (slab_half bytes
    // further) by the lo slab of the same plane; per tap and k-step the issuer adds x_lo * W_hi (an N = cout_pad MMA on
    // the first cout_pad rows of the dual weight tile, idesc2) to x_hi * [W_hi | W_lo].  The input is read once instead
    // of twice and the output is written once instead of write + read-modify-write.
    int tri; uint32_t slab_half, idesc2;
    uint32_t idesc;
    uint32_t slab_bytes, w_bytes;
    int debug;                    // dev only (LFB200_TC_DEBUG): 1 skip MMAs, 2 skip producer work, 4 skip epilogue work
    uint64_t magic_q4, magic_P;   // ceil(2^40 / divisor): exact n / divisor for n < 2^20, divisor < 2^12
};

__device__ __forceinline__ int fast_div(int n, uint64_t magic) { return (int)(((uint64_t)(uint32_t)n * magic) >> 40); }

// dev-only timeline of CTA 0 (LFB200_TC_DEBUG & 8): [role][event][2] SM clock stamps
__device__ long long g_dbg[3][64][2];
__device__ __forceinline__ void dbg_stamp(const Params& p, int role, int idx, int which) {
    if ((p.debug & 8) && blockIdx.x == 0 && idx < 64) g_dbg[role][idx][which] = clock64();
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t it = 0; it < (1u << 26); ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    printf("lfb200 conv_tc: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n",
           (int)blockIdx.x, (int)threadIdx.x, bar, parity);
    __trap();
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine (UBLKCP), completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> f32, M=128, N from idesc, K=16
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// same, with the two descriptors given as (lo, hi) 32-bit halves (cheap to update in the issue loop)
__device__ __forceinline__ void umma_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}

// UMMA shared-memory matrix descriptor, no-swizzle K-major canonical layout
// (cute::UMMA::SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout 0)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));    // one instruction for two values
    return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b, int part) {
    uint32_t h = cvt_bf16x2(a, b);
    if (part) {      // residual: x - bf16(x), itself rounded to bf16
        const float ra = a - __uint_as_float(h << 16), rb = b - __uint_as_float(h & 0xffff0000u);
        h = cvt_bf16x2(ra, rb);
    }
    return h;
}

// One output position per thread: its Cout accumulators come out of TMEM into registers, so the
// Equalized/Block epilogue (He scale, bias, LeakyReLU, PixelNorm over channels) is thread-local.
// NCH = Cout_pad/16 known at compile time (registers hold the whole row); NCH == 0 is the generic
// two-pass variant for wide layers.  Branch-free: bias is staged zero-padded in smem, LeakyReLU is
// max(a, slope*a) with slope folded to 1 when disabled, PixelNorm multiplies by a per-row reciprocal.
template <int NCH, int MODE>
__device__ __forceinline__ void epilogue_tile(const Params& p, const float* __restrict__ bias_s, uint32_t taddr,
                                              int64_t opos, bool valid) {
    constexpr bool kRaw = (MODE == PASS_FIRST || MODE == PASS_MID);
    constexpr bool kAdd = (MODE == PASS_MID || MODE == PASS_LAST);
    float* yp = p.y + opos * p.cout;
    const float slope = p.act ? p.slope : 1.f;
    if (NCH > 0) {
        float v[NCH > 0 ? NCH * 16 : 16];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            float t16[16];
            tmem_ld16(taddr + ch * 16, t16);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[ch * 16 + i] = t16[i];
        }
        if (p.dual) {                                  // second half of the accumulator row: x_hi * W_lo
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                float t16[16];
                tmem_ld16(taddr + p.cout_pad + ch * 16, t16);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[ch * 16 + i] += t16[i];
            }
        }
        if (valid) {
            if (kAdd) {
#pragma unroll
                for (int i = 0; i < NCH * 16; i += 4) {
                    if (i < p.cout) {
                        const float4 o = *reinterpret_cast<const float4*>(yp + i);
                        v[i] += o.x; v[i + 1] += o.y; v[i + 2] += o.z; v[i + 3] += o.w;
                    }
                }
            }
            if (!kRaw) {
                float ss = 0.f;
#pragma unroll
                for (int i = 0; i < NCH * 16; ++i) {
                    float a = v[i] * p.scale + bias_s[i];
                    a = fmaxf(a, a * slope);
                    v[i] = a;
                    ss += a * a;
                }
                if (p.norm) {
                    const float rn = sqrtf(ss / (float)p.cout + 1e-8f);
                    const float inv = 1.f / rn;
#pragma unroll
                    for (int i = 0; i < NCH * 16; ++i) v[i] *= inv;
                    if (p.rnorm != nullptr) p.rnorm[opos] = rn;
                }
                if (p.epi_y != nullptr) {
                    // du_prev = gate(y) * (g - y * mean_c(g*y)) / r   with g = this row (same formula as lf_actnorm_bwd)
                    const float* yr = p.epi_y + opos * p.cout;
                    float dot = 0.f, ir = 1.f;
                    if (p.epi_norm) {
#pragma unroll
                        for (int i = 0; i < NCH * 16; i += 4) {
                            if (i < p.cout) {
                                const float4 y4 = ldg4(yr + i);
                                dot += v[i] * y4.x + v[i + 1] * y4.y + v[i + 2] * y4.z + v[i + 3] * y4.w;
                            }
                        }
                        dot *= 1.f / (float)p.cout;
                        ir = 1.f / __ldg(p.epi_r + opos);
                    }
                    const float gs = p.epi_act ? p.epi_slope : 1.f;
#pragma unroll
                    for (int i = 0; i < NCH * 16; i += 4) {
                        if (i < p.cout) {
                            const float4 y4 = ldg4(yr + i);
                            const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float o = (v[i + j] - yv[j] * dot) * ir;
                                v[i + j] = yv[j] > 0.f ? o : o * gs;
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NCH * 16; i += 4)
                if (i < p.cout) *reinterpret_cast<float4*>(yp + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
        __syncwarp();
    } else {
        float ss = 0.f;
        if (p.norm && !kRaw) {
            for (int cb = 0; cb < p.cout_pad; cb += 16) {
                float v[16];
                __syncwarp();
                tmem_ld16(taddr + cb, v);
                if (p.dual) {
                    float u[16];
                    tmem_ld16(taddr + p.cout_pad + cb, u);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += u[i];
                }
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float a = v[i];
                        if (kAdd && cb + i < p.cout) a += yp[cb + i];
                        a = a * p.scale + bias_s[cb + i];
                        a = fmaxf(a, a * slope);
                        ss += a * a;
                    }
                }
            }
        }
        const float rn = sqrtf(ss / (float)p.cout + 1e-8f);
        const float inv = (p.norm && !kRaw) ? 1.f / rn : 1.f;
        for (int cb = 0; cb < p.cout_pad; cb += 16) {
            float v[16];
            __syncwarp();
            tmem_ld16(taddr + cb, v);
            if (p.dual) {
                float u[16];
                tmem_ld16(taddr + p.cout_pad + cb, u);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += u[i];
            }
            if (valid) {
                if ((p.cout & 3) == 0) {                 // 128-bit read-modify-write of the row
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        if (cb + i < p.cout) {
                            float a[4] = {v[i], v[i + 1], v[i + 2], v[i + 3]};
                            if (kAdd) {
                                const float4 o = *reinterpret_cast<const float4*>(yp + cb + i);
                                a[0] += o.x; a[1] += o.y; a[2] += o.z; a[3] += o.w;
                            }
                            if (!kRaw) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    a[j] = a[j] * p.scale + bias_s[cb + i + j];
                                    a[j] = fmaxf(a[j], a[j] * slope) * inv;
                                }
                            }
                            *reinterpret_cast<float4*>(yp + cb + i) = make_float4(a[0], a[1], a[2], a[3]);
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float a = v[i];
                        if (cb + i < p.cout) {
                            if (kAdd) a += yp[cb + i];
                            if (!kRaw) {
                                a = a * p.scale + bias_s[cb + i];
                                a = fmaxf(a, a * slope) * inv;
                            }
                            yp[cb + i] = a;
                        }
                    }
                }
            }
        }
        if (valid && p.norm && !kRaw && p.rnorm != nullptr) p.rnorm[opos] = rn;
        __syncwarp();
    }
}

template <int NCH>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    const uint32_t slabs = smem_u32(smem);                              // ring * slab_bytes
    const uint32_t wsm = slabs + p.ring * p.slab_bytes;                 // packed bf16 weights
    uint8_t* tail = smem + (size_t)p.ring * p.slab_bytes + p.w_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);                 // [ring] full, [ring] empty, [2] acc_full, [2] acc_empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxRing + 4);
    float* bias_s = reinterpret_cast<float*>(bars + 2 * kMaxRing + 6);     // [cout_pad], zero padded
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kMaxRing);
    const uint32_t bar_accf = smem_u32(bars + 2 * kMaxRing), bar_acce = smem_u32(bars + 2 * kMaxRing + 2);
    const uint32_t bar_w = smem_u32(bars + 2 * kMaxRing + 5);           // packed weights have landed (bulk TMA)

    // warp index broadcast from lane 0 so the compiler knows it is warp-uniform: everything derived from it
    // (tile id, TMEM/descriptor addresses) can then live in uniform registers, which UTCHMMA consumes directly
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    // ---- one-time setup: barriers, TMEM, bias; the packed weights come in by bulk TMA while the producers already
    //      stage the first slab (the issuers wait on bar_w before their first MMA)
    for (int i = threadIdx.x; i < p.cout_pad; i += kThreads)
        bias_s[i] = (p.bias != nullptr && i < p.cout) ? p.bias[i] : 0.f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.ring; ++i) { mbar_init(bar_full + 8 * i, p.nprod * 32); mbar_init(bar_empty + 8 * i, p.nmma); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_accf + 8 * i, p.nmma); mbar_init(bar_acce + 8 * i, 32 * kEpiWarps); }
        mbar_init(bar_w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arrive_expect_tx(bar_w, p.w_bytes);
        for (uint32_t off = 0; off < p.w_bytes; off += 32768u) {
            const uint32_t nb = p.w_bytes - off < 32768u ? p.w_bytes - off : 32768u;
            bulk_g2s(wsm + off, reinterpret_cast<const uint8_t*>(p.wpk) + off, nb, bar_w);
        }
    }
    const int NPROD = p.nprod, NMMA = p.nmma;
    if (warp == NPROD) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int halo = p.k / 2;
    const int rows_in = p.R + 2 * halo;
    const int kchunks = p.cin_pad / 8;
    const uint32_t lbo_a = (uint32_t)p.pos_alloc * 16u;
    const int q4 = p.cin_pad / 4;                       // float4 units per position

    if (warp < NPROD) {
        // =========================== PRODUCERS: global fp32 -> bf16 UMMA slab ===========================
        uint32_t kcount = 0;                            // planes produced by this CTA so far
        const int tid = threadIdx.x;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            int it = item;
            const int dchunk = it % p.ndchunks; it /= p.ndchunks;
            const int strip = it % p.nstrips; const int n = it / p.nstrips;
            const int d0 = dchunk * p.DC, d1 = min(p.d, d0 + p.DC);
            const int y0 = strip * p.R;
            for (int e = d0 - p.hz; e <= d1 - 1 + p.hz; ++e, ++kcount) {
                const uint32_t slot = kcount % p.ring;
                mbar_wait(bar_empty + 8 * slot, ((kcount / p.ring) & 1) ^ 1);
                if (tid == 0) dbg_stamp(p, 0, kcount, 0);
                if (e >= 0 && e < p.d && !(p.debug & 2)) {
                    uint8_t* slab = smem + (size_t)slot * p.slab_bytes;
                    const float* plane = p.x + ((int64_t)n * p.d + e) * p.h * p.w * (int64_t)p.cin;
                    const int units = rows_in * p.P * q4;
                    if (p.pro_y == nullptr) {
                    constexpr int kBatch = 8;           // loads in flight per thread
                    for (int u0 = tid; u0 < units; u0 += kBatch * NPROD * 32) {
                        float4 v[kBatch];
                        int udst[kBatch];               // byte offset inside the slab, -1 = nothing to do
#pragma unroll
                        for (int j = 0; j < kBatch; ++j) {
                            const int u = u0 + j * NPROD * 32;
                            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                            udst[j] = -1;
                            if (u < units) {
                                const int pos = fast_div(u, p.magic_q4), q = u - pos * q4;
                                const int r = fast_div(pos, p.magic_P), c = pos - r * p.P;
                                const int yy = y0 + r - halo, xx = c - halo;
                                udst[j] = (q >> 1) * (int)lbo_a + pos * 16 + (q & 1) * 8;
                                if (yy >= 0 && yy < p.h && xx >= 0 && xx < p.w && q * 4 < p.cin)
                                    v[j] = ldg4(plane + ((int64_t)yy * p.w + xx) * p.cin + q * 4);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < kBatch; ++j) {
                            if (udst[j] < 0) continue;
                            if (p.tri) {
                                const uint32_t h0 = cvt_bf16x2(v[j].x, v[j].y), h1 = cvt_bf16x2(v[j].z, v[j].w);
                                const uint32_t l0 = cvt_bf16x2(v[j].x - __uint_as_float(h0 << 16), v[j].y - __uint_as_float(h0 & 0xffff0000u));
                                const uint32_t l1 = cvt_bf16x2(v[j].z - __uint_as_float(h1 << 16), v[j].w - __uint_as_float(h1 & 0xffff0000u));
                                *reinterpret_cast<uint2*>(slab + udst[j]) = make_uint2(h0, h1);
                                *reinterpret_cast<uint2*>(slab + p.slab_half + udst[j]) = make_uint2(l0, l1);
                                continue;
                            }
                            const uint32_t lo = pack_bf16x2(v[j].x, v[j].y, p.a_part);
                            const uint32_t hi = pack_bf16x2(v[j].z, v[j].w, p.a_part);
                            *reinterpret_cast<uint2*>(slab + udst[j]) = make_uint2(lo, hi);
                        }
                    }
                    } else {
                        // ---- fused PixelNorm/LeakyReLU backward: du = gate(y) * (g - y * mean_c(g*y)) / r
                        // a position's channels sit on 2^pro_lg consecutive lanes (q4 is a power of two that
                        // divides 32 and the producer thread count), so mean_c is an xor-shuffle reduction
                        const float* yplane = p.pro_y + ((int64_t)n * p.d + e) * p.h * p.w * (int64_t)p.cin;
                        const float* rplane = p.pro_r + ((int64_t)n * p.d + e) * p.h * p.w;
                        const float inv_c = 1.f / (float)p.cin;
                        const int units_pad = (units + 31) & ~31;
                        constexpr int kB2 = 4;
                        for (int u0 = tid; u0 < units_pad; u0 += kB2 * NPROD * 32) {
                            float4 g4[kB2], y4[kB2];
                            float rr[kB2];
                            int udst[kB2];
#pragma unroll
                            for (int j = 0; j < kB2; ++j) {
                                const int u = u0 + j * NPROD * 32;
                                g4[j] = make_float4(0.f, 0.f, 0.f, 0.f); y4[j] = g4[j]; rr[j] = 1.f; udst[j] = -1;
                                if (u < units) {
                                    const int pos = fast_div(u, p.magic_q4), q = u - pos * q4;
                                    const int r = fast_div(pos, p.magic_P), c = pos - r * p.P;
                                    const int yy = y0 + r - halo, xx = c - halo;
                                    udst[j] = (q >> 1) * (int)lbo_a + pos * 16 + (q & 1) * 8;
                                    if (yy >= 0 && yy < p.h && xx >= 0 && xx < p.w) {
                                        const int64_t gp = (int64_t)yy * p.w + xx;
                                        g4[j] = ldg4(plane + gp * p.cin + q * 4);
                                        y4[j] = ldg4(yplane + gp * p.cin + q * 4);
                                        if (p.pro_norm) rr[j] = __ldg(rplane + gp);
                                    }
                                }
                            }
#pragma unroll
                            for (int j = 0; j < kB2; ++j) {
                                float4 g = g4[j];
                                const float4 yv = y4[j];
                                if (p.pro_norm) {
                                    float dot = g.x * yv.x + g.y * yv.y + g.z * yv.z + g.w * yv.w;
                                    for (int o = 1; o < (1 << p.pro_lg); o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                                    dot *= inv_c;
                                    const float ir = 1.f / rr[j];
                                    g.x = (g.x - yv.x * dot) * ir; g.y = (g.y - yv.y * dot) * ir;
                                    g.z = (g.z - yv.z * dot) * ir; g.w = (g.w - yv.w * dot) * ir;
                                }
                                if (p.pro_act) {
                                    g.x = yv.x > 0.f ? g.x : g.x * p.pro_slope; g.y = yv.y > 0.f ? g.y : g.y * p.pro_slope;
                                    g.z = yv.z > 0.f ? g.z : g.z * p.pro_slope; g.w = yv.w > 0.f ? g.w : g.w * p.pro_slope;
                                }
                                if (udst[j] >= 0) {
                                    const uint32_t lo = pack_bf16x2(g.x, g.y, p.a_part);
                                    const uint32_t hi = pack_bf16x2(g.z, g.w, p.a_part);
                                    *reinterpret_cast<uint2*>(slab + udst[j]) = make_uint2(lo, hi);
                                }
                            }
                        }
                    }
                    fence_proxy_async();                // generic-proxy stores -> visible to the tensor core
                }
                mbar_arrive(bar_full + 8 * slot);
                if (tid == 0) dbg_stamp(p, 0, kcount, 1);
            }
        }
    } else if (warp < NPROD + NMMA) {
        // =========================== MMA ISSUERS (warp m owns M-tile m) ===========================
        // The whole warp runs this loop in lock-step so every descriptor/address lives in UNIFORM registers
        // (UTCHMMA takes uniform operands; a lane-0-only branch makes the compiler shuttle them through
        // R2UR + an elect waterfall, ~12 instructions per MMA); only the issue itself is elected.
        const int my_tile = warp - NPROD;
        if (my_tile < p.NT) {
            uint32_t kbase = 0;                         // running plane count at the start of the item
            uint32_t step = 0;                          // running step (accumulator) count
            const uint32_t lbo_b = (uint32_t)p.ncols * 16u;
            // descriptor pieces (16-byte units): lo = start | LBO << 16 ; hi = SBO(128 B) | version 1 << 14
            const uint32_t desc_hi = (128u >> 4) | (1u << 14);
            const uint32_t a_lo_const = (lbo_a >> 4) << 16, b_lo_const = (lbo_b >> 4) << 16;
            const uint32_t a_kstride = 2 * (lbo_a >> 4), b_kstride = 2 * (lbo_b >> 4);
            const uint32_t tap_stride = (uint32_t)kchunks * (lbo_b >> 4);
            const int ksteps = p.cin_pad / 16;
            // hot-loop operands copied out of the (constant-bank) parameter block: the "memory" clobber of the
            // MMA asm would otherwise force a constant reload (LDCU, ~40 cycles) on every loop test
            const int K = p.k, HZ = p.hz, NT = p.NT, RING = p.ring, DEPTH = p.d, DBG = p.debug;
            const uint32_t PP = (uint32_t)p.P, IDESC = p.idesc, COUT_PAD = (uint32_t)p.ncols, SLAB = p.slab_bytes;
            const uint32_t TRI = (uint32_t)p.tri, IDESC2 = p.idesc2, HALF16 = p.slab_half >> 4;
            mbar_wait(bar_w, 0);
            for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
                const int dchunk = item % p.ndchunks;
                const int d0 = dchunk * p.DC, d1 = min(p.d, d0 + p.DC);
                const int nplanes = (d1 - d0) + 2 * p.hz;
                int waited = 0;                         // planes of this item whose full barrier we passed
                for (int d = d0; d < d1; ++d, ++step) {
                    const int need = (d - d0) + 2 * p.hz + 1;        // planes 0 .. need-1 must be staged
                    for (; waited < need; ++waited) {
                        const uint32_t kc = kbase + waited;
                        mbar_wait(bar_full + 8 * (kc % p.ring), (kc / p.ring) & 1);
                    }
                    const uint32_t buf = step & 1;
                    mbar_wait(bar_acce + 8 * buf, ((step >> 1) & 1) ^ 1);
                    tc_fence_after();
                    if (my_tile == 0 && lane == 0) dbg_stamp(p, 1, step, 0);
                    // Issue loop kept to a handful of instructions per MMA: the descriptor's high word is
                    // constant, the low word is (LBO field | start>>4) and only the start changes, by +1 per
                    // dx, +P per dy, +2*LBO per k-step (all in 16-byte units).
                    for (int t = my_tile; t < NT; t += NMMA) {
                        const uint32_t d_tmem = tmem_base + (buf * NT + t) * COUT_PAD;
                        uint32_t acc = 0;
                        for (int dz = 0; dz <= 2 * HZ; ++dz) {
                            const int e = d + dz - HZ;
                            if (e < 0 || e >= DEPTH) continue;         // zero plane: contributes nothing
                            const uint32_t kc = kbase + (uint32_t)(e - (d0 - HZ));
                            uint32_t a_row = a_lo_const | (((slabs + (kc % RING) * SLAB) >> 4) + (uint32_t)t * 128u);
                            uint32_t b_cur = b_lo_const | ((wsm >> 4) + (uint32_t)(dz * K * K) * tap_stride);
                            for (int dy = 0; dy < K; ++dy) {
                                uint32_t a_cur = a_row;
                                for (int dx = 0; dx < K; ++dx) {
                                    uint32_t ak = a_cur, bk = b_cur;
                                    for (int ks = 0; ks < ksteps; ++ks) {
                                        if (!(DBG & 1) && elect_one()) {
                                            umma_bf16_lohi(d_tmem, ak, desc_hi, bk, desc_hi, IDESC, acc);
                                            if (TRI) umma_bf16_lohi(d_tmem, ak + HALF16, desc_hi, bk, desc_hi, IDESC2, 1u);
                                        }
                                        acc = 1;
                                        ak += a_kstride; bk += b_kstride;
                                    }
                                    a_cur += 1;
                                    b_cur += tap_stride;
                                }
                                a_row += PP;
                            }
                        }
                    }
                    if (elect_one()) umma_commit(bar_accf + 8 * buf);  // accumulators of this step complete
                    if (my_tile == 0 && lane == 0) dbg_stamp(p, 1, step, 1);
                    // release planes that no later step of this item reads
                    const int first_rel = (d - d0);                    // plane index d - hz relative to item
                    const int last_rel = (d == d1 - 1) ? nplanes - 1 : first_rel;
                    for (int pl = first_rel; pl <= last_rel; ++pl)
                        if (elect_one()) umma_commit(bar_empty + 8 * ((kbase + pl) % p.ring));
                    __syncwarp();
                }
                kbase += nplanes;
            }
        }
    } else {
        // =========================== EPILOGUE (4 warps = 128 TMEM lanes) ===========================
        const int wq = warp & 3;                       // TMEM lane quarter this warp may access
        const int ehalf = (warp - NPROD - NMMA) >> 2;   // which half of the tiles this warp drains (0 with 4 epilogue warps)
        uint32_t step = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            int it = item;
            const int dchunk = it % p.ndchunks; it /= p.ndchunks;
            const int strip = it % p.nstrips; const int n = it / p.nstrips;
            const int d0 = dchunk * p.DC, d1 = min(p.d, d0 + p.DC);
            const int y0 = strip * p.R;
            for (int d = d0; d < d1; ++d, ++step) {
                const uint32_t buf = step & 1;
                mbar_wait(bar_accf + 8 * buf, (step >> 1) & 1);
                tc_fence_after();
                if (wq == 0 && ehalf == 0 && lane == 0) dbg_stamp(p, 2, step, 0);
                for (int t = ehalf; t < ((p.debug & 4) ? 0 : p.NT); t += kEpiWarps / 4) {
                    const int q = t * 128 + wq * 32 + lane;
                    const int r = fast_div(q, p.magic_P), c = q - r * p.P;
                    const bool valid = (r < p.R) && (c < p.w) && (y0 + r < p.h);
                    const int64_t opos = (((int64_t)n * p.d + d) * p.h + (y0 + r)) * p.w + c;
                    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (buf * p.NT + t) * p.ncols;
                    switch (p.pass_mode) {
                        case PASS_ONLY:  epilogue_tile<NCH, PASS_ONLY>(p, bias_s, taddr, opos, valid); break;
                        case PASS_FIRST: epilogue_tile<NCH, PASS_FIRST>(p, bias_s, taddr, opos, valid); break;
                        case PASS_MID:   epilogue_tile<NCH, PASS_MID>(p, bias_s, taddr, opos, valid); break;
                        default:         epilogue_tile<NCH, PASS_LAST>(p, bias_s, taddr, opos, valid); break;
                    }
                }
                tc_fence_before();
                mbar_arrive(bar_acce + 8 * buf);
                if (wq == 0 && ehalf == 0 && lane == 0) dbg_stamp(p, 2, step, 1);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == NPROD) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// fp32 packed weights [taps][cin][cout] -> bf16 UMMA layout [part(hi,lo)][tap][k-chunk][cout_pad][8]
__global__ void pack_weights_kernel(const float* __restrict__ w, uint16_t* __restrict__ out,
                                    int taps, int cin, int cout, int cin_pad, int cout_pad) {
    const int64_t per_part = (int64_t)taps * (cin_pad / 8) * cout_pad * 8;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < per_part; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = e;
        const int j = (int)(r % 8); r /= 8;
        const int co = (int)(r % cout_pad); r /= cout_pad;
        const int kc = (int)(r % (cin_pad / 8)); const int tap = (int)(r / (cin_pad / 8));
        const int ci = kc * 8 + j;
        float v = 0.f;
        if (ci < cin && co < cout) v = w[((int64_t)tap * cin + ci) * cout + co];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        out[e] = __bfloat16_as_ushort(hi);
        out[per_part + e] = __bfloat16_as_ushort(lo);
        // dual layout [tap][k-chunk][2*cout_pad][8]: rows [0,cout_pad) = hi, [cout_pad, 2*cout_pad) = lo
        const int64_t drow = ((int64_t)tap * (cin_pad / 8) + kc) * (2 * cout_pad);
        out[2 * per_part + (drow + co) * 8 + j] = __bfloat16_as_ushort(hi);
        out[2 * per_part + (drow + cout_pad + co) * 8 + j] = __bfloat16_as_ushort(lo);
    }
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

struct Plan {
    int cin_pad, cout_pad, P, R, NT, pos_alloc, ring, DC, nstrips, ndchunks, taps;
    uint32_t slab_bytes, w_bytes, smem_bytes, slab_half;
};

static bool make_plan(const lf_conv_desc* d, Plan& pl, bool dual = false, bool tri = false) {
    if (!(d->ndim == 2 || d->ndim == 3)) return false;
    if (!(d->k == 1 || d->k == 3)) return false;
    if (d->cin % 4 != 0) return false;
    pl.cin_pad = round_up(d->cin, 16);
    pl.cout_pad = round_up(d->cout, 16);
    if (pl.cout_pad > 256) return false;
    const int halo = d->k / 2;
    const int hz = (d->ndim == 3) ? halo : 0;
    pl.taps = (d->ndim == 3) ? d->k * d->k * d->k : d->k * d->k;
    pl.P = d->w + 2 * halo;
    if (pl.P >= 4096 || pl.cin_pad > 1024) return false;
    // accumulators: 2 buffers x NT tiles x cout_pad columns <= 512
    const int ncols = pl.cout_pad * (dual ? 2 : 1);
    if (ncols > 256) return false;
    int nt_max = 512 / (2 * ncols);
    if (nt_max > kMaxTiles) nt_max = kMaxTiles;
    if (nt_max < 1) return false;
    pl.w_bytes = (uint32_t)pl.taps * (pl.cin_pad / 8) * ncols * 16;
    const uint32_t budget = 227 * 1024 - 256 - 1024;
    if (pl.w_bytes + 4096 > budget) return false;
    pl.ring = (hz > 0) ? kMaxRing : 2;
    // largest R such that tiles and smem fit
    int best_R = 0;
    for (int R = min(d->h, (nt_max * 128) / pl.P); R >= 1; --R) {
        int pos = (R + 2 * halo) * pl.P + 2 * halo;
        pos = round_up(pos - 4, 8) + 4;                 // pos_alloc == 4 (mod 8): conflict-free producer stores
        const uint32_t slab = (uint32_t)(pl.cin_pad / 8) * pos * 16 * (tri ? 2u : 1u);
        if ((uint64_t)slab * pl.ring + pl.w_bytes <= budget && ((uint32_t)pos * 16 >> 4) < 16384) { best_R = R; break; }
    }
    if (best_R == 0) return false;
    if (hz == 0) {
        // 2-D layers have no depth to chunk; the strip height trades halo re-reads and per-item latency against the
        // number of rounds the persistent CTAs make.  Cost model (in staged positions): rounds x (staged positions of
        // one item + ~800 for its fixed fill / MMA / drain latency).  E.g. 8 maps of 64^2: R = 2 gave 256 items = two
        // rounds on 148 SMs with half-empty second M-tiles; R = 4 is 128 items, one round.
        const int sms = sm_count();
        int64_t best_cost = -1;
        int pick = best_R;
        for (int R = best_R; R >= (best_R < 2 ? best_R : 2); --R) {
            const int64_t items = (int64_t)d->n * ((d->h + R - 1) / R);
            const int64_t rounds = (items + sms - 1) / sms;
            const int64_t cost = rounds * ((int64_t)(R + 2 * halo) * pl.P + 800);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; pick = R; }
        }
        best_R = pick;
    }
    pl.R = best_R;
    pl.NT = (pl.R * pl.P + 127) / 128;
    int pos = (pl.R + 2 * halo) * pl.P + 2 * halo;
    pl.pos_alloc = round_up(pos - 4, 8) + 4;
    pl.slab_half = (uint32_t)(pl.cin_pad / 8) * pl.pos_alloc * 16;
    pl.slab_bytes = pl.slab_half * (tri ? 2u : 1u);
    pl.smem_bytes = pl.slab_bytes * pl.ring + pl.w_bytes + 256 + 1024;
    // the last tile may read up to (NT*128 + (k-1)*P + k-1 - pos_alloc) positions past a k-chunk: it must stay
    // inside the allocation; chunks are followed by other chunks / the weight buffer, check the very last one
    const int overrun = pl.NT * 128 + (d->k - 1) * pl.P + (d->k - 1) - pl.pos_alloc;
    if (overrun > 0 && (uint32_t)overrun * 16 > pl.w_bytes) return false;
    pl.nstrips = (d->h + pl.R - 1) / pl.R;
    // depth chunking: enough work items to fill the machine
    pl.DC = d->d;
    if (hz > 0) {
        const int sms = sm_count();
        // enough items for one balanced wave; longer depth chunks amortise the 2 halo planes (measured:
        // DC 8 -> 0.33 ms, 16 -> 0.31, 32 -> 0.30 for the config-B 3x3x3 conv)
        while (pl.DC > 4 && (int64_t)d->n * pl.nstrips * ((d->d + pl.DC - 1) / pl.DC) < sms) pl.DC = (pl.DC + 1) / 2;
    }
    { const int dc = option(OPT_TC_DC); if (hz > 0 && dc > 0) pl.DC = min(d->d, dc); }
    pl.ndchunks = (d->d + pl.DC - 1) / pl.DC;
    return true;
}

}  // namespace tc

int conv_tc_supported(const lf_conv_desc* d) {
    tc::Plan pl;
    return tc::make_plan(d, pl) ? 1 : 0;
}

struct TcPrologue { const float* y; const float* rnorm; int act, norm; float slope; };
struct TcEpilogue { const float* y; const float* rnorm; int act, norm; float slope; };

int conv_tc_launch_ex(const lf_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                      float* rnorm, const TcPrologue* pro, const TcEpilogue* epi, cudaStream_t st);

int conv_tc_launch(const lf_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                   float* rnorm, cudaStream_t st) {
    return conv_tc_launch_ex(d, x, w, bias, y, rnorm, nullptr, nullptr, st);
}

// one kernel launch: `wpk` points at the packed weight region to use (hi, lo or dual)
static int conv_tc_launch_pass(const lf_conv_desc* d, const tc::Plan& pl, const float* x, const uint16_t* wpk,
                               const float* bias, float* y, float* rnorm, const TcPrologue* pro, int a_part,
                               int mode, int dual, cudaStream_t st, const TcEpilogue* epi = nullptr, int tri = 0) {
    tc::Params p;
    p.tri = tri; p.slab_half = pl.slab_half;
    p.idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(pl.cout_pad >> 3) << 17) | ((128u >> 4) << 24);
    p.epi_y = epi ? epi->y : nullptr; p.epi_r = epi ? epi->rnorm : nullptr;
    p.epi_act = epi ? epi->act : 0; p.epi_norm = epi ? epi->norm : 0; p.epi_slope = epi ? epi->slope : 1.f;
    if (epi) LF_CHECK_ARG((d->cout & 3) == 0 && pl.cout_pad <= 32 && (mode == tc::PASS_ONLY || mode == tc::PASS_LAST),
                          "conv_tc: fused backward epilogue needs Cout % 4 == 0, Cout <= 32, on the final pass");
    p.x = x; p.bias = bias; p.y = y; p.rnorm = rnorm; p.wpk = wpk;
    p.a_part = a_part; p.pass_mode = mode; p.dual = dual;
    p.pro_y = pro ? pro->y : nullptr; p.pro_r = pro ? pro->rnorm : nullptr;
    p.pro_act = pro ? pro->act : 0; p.pro_norm = pro ? pro->norm : 0; p.pro_slope = pro ? pro->slope : 1.f; p.pro_lg = 0;
    if (pro) {
        const int q4 = pl.cin_pad / 4;
        LF_CHECK_ARG(d->cin == pl.cin_pad && (q4 & (q4 - 1)) == 0 && q4 <= 32, "conv_tc: fused backward prologue needs Cin in {16,32,64,128}");
        while ((1 << p.pro_lg) < q4) ++p.pro_lg;
    }
    p.n = d->n; p.d = d->d; p.h = d->h; p.w = d->w;
    p.cin = d->cin; p.cout = d->cout; p.cin_pad = pl.cin_pad; p.cout_pad = pl.cout_pad;
    p.ncols = pl.cout_pad * (dual ? 2 : 1);
    p.nmma = pl.NT < tc::kMmaWarps ? pl.NT : tc::kMmaWarps;
    p.nprod = tc::kWarps - tc::kEpiWarps - p.nmma;
    p.k = d->k; p.hz = (d->ndim == 3) ? d->k / 2 : 0;
    p.R = pl.R; p.NT = pl.NT; p.P = pl.P; p.pos_alloc = pl.pos_alloc; p.ring = pl.ring; p.DC = pl.DC;
    p.nstrips = pl.nstrips; p.ndchunks = pl.ndchunks; p.items = d->n * pl.nstrips * pl.ndchunks;
    p.scale = d->scale; p.act = d->act; p.slope = d->slope; p.norm = d->norm;
    p.slab_bytes = pl.slab_bytes; p.w_bytes = pl.w_bytes;
    p.debug = option(OPT_TC_DEBUG);
    p.magic_q4 = ((1ull << 40) + (pl.cin_pad / 4) - 1) / (uint64_t)(pl.cin_pad / 4);
    p.magic_P = ((1ull << 40) + pl.P - 1) / (uint64_t)pl.P;
    // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=BF16 [7,10)=1, b=BF16 [10,13)=1,
    // K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.ncols >> 3) << 17) | ((128u >> 4) << 24);
    void (*kern)(tc::Params) = nullptr;
    switch (((d->cout & 3) == 0 && pl.cout_pad <= 32) ? pl.cout_pad / 16 : 0) {
        case 1: kern = tc::conv_tc_kernel<1>; break;
        case 2: kern = tc::conv_tc_kernel<2>; break;
        default: kern = tc::conv_tc_kernel<0>; break;
    }
    {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("conv_tc: cannot raise dynamic smem: %s", cudaGetErrorString(e)); return (int)e; }
    }
    const int grid = min(p.items, sm_count());
    kern<<<grid, tc::kThreads, pl.smem_bytes, st>>>(p);
    LF_RETURN_LAUNCH();
}

int conv_tc_launch_ex(const lf_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                      float* rnorm, const TcPrologue* pro, const TcEpilogue* epi, cudaStream_t st) {
    tc::Plan pl;
    LF_CHECK_ARG(tc::make_plan(d, pl), "conv_tc: unsupported shape");
    LF_CHECK_ARG(x && w && y, "conv_tc: null pointer");
    // packed weights: [hi | lo | dual], each `part` elements (dual is two parts long)
    const uint16_t* wbase = reinterpret_cast<const uint16_t*>(w);
    const size_t part = (size_t)pl.w_bytes / 2;
    if (d->precision == 2) return conv_tc_launch_pass(d, pl, x, wbase, bias, y, rnorm, pro, 0, tc::PASS_ONLY, 0, st, epi);
    const bool no_dual = option(OPT_TC_NO_DUAL) != 0;
    if (!no_dual && option(OPT_TC_NO_TRI) == 0 && pro == nullptr && d->ndim == 2 && (d->cout & 3) == 0 && pl.cout_pad <= 64) {
        // 2-D layers: the whole bf16x3 product in one launch (hi and lo slabs staged side by side)
        tc::Plan pt;
        if (tc::make_plan(d, pt, true, true))
            return conv_tc_launch_pass(d, pt, x, wbase + 2 * part, bias, y, rnorm, nullptr, 0, tc::PASS_ONLY, 1, st, epi, 1);
    }
    if (!no_dual && pro == nullptr && (d->cout & 3) == 0 && pl.cout_pad <= 64) {
        // bf16x3 in TWO passes: x_hi * [W_hi | W_lo] (one N = 2*Cout MMA per tap: the A tile is fetched from shared
        // memory once for both products), then x_lo * W_hi accumulated in the epilogue
        tc::Plan pd;
        if (tc::make_plan(d, pd, true)) {
            int e = conv_tc_launch_pass(d, pd, x, wbase + 2 * part, bias, y, rnorm, nullptr, 0, tc::PASS_FIRST, 1, st);
            if (e != LF_OK) return e;
            if (option(OPT_TC_DEBUG) & 16) return LF_OK;   // profiling: first pass only
            return conv_tc_launch_pass(d, pl, x, wbase, bias, y, rnorm, nullptr, 1, tc::PASS_LAST, 0, st, epi);
        }
    }
    int e = conv_tc_launch_pass(d, pl, x, wbase, bias, y, rnorm, pro, 0, tc::PASS_FIRST, 0, st);
    if (e != LF_OK) return e;
    e = conv_tc_launch_pass(d, pl, x, wbase, bias, y, rnorm, pro, 1, tc::PASS_MID, 0, st);
    if (e != LF_OK) return e;
    return conv_tc_launch_pass(d, pl, x, wbase + part, bias, y, rnorm, pro, 0, tc::PASS_LAST, 0, st, epi);
}

}  // namespace lf

using namespace lf;

extern "C" int64_t lf_conv_tc_weight_bytes(int taps, int cin, int cout) {
    if (taps <= 0 || cin <= 0 || cout <= 0) return 0;
    const int64_t cin_pad = (cin + 15) / 16 * 16, cout_pad = (cout + 15) / 16 * 16;
    return 4 * taps * cin_pad * cout_pad * 2;       // hi | lo | dual (hi and lo interleaved per tap / k-chunk)
}

extern "C" int lf_conv_tc_pack_weights(const float* w_packed, void* out, int taps, int cin, int cout, void* stream) {
    LF_CHECK_ARG(w_packed && out && taps > 0 && cin > 0 && cout > 0, "conv_tc_pack_weights: bad arguments");
    const int cin_pad = (cin + 15) / 16 * 16, cout_pad = (cout + 15) / 16 * 16;
    const int64_t per_part = (int64_t)taps * cin_pad * cout_pad;
    tc::pack_weights_kernel<<<(unsigned)((per_part + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        w_packed, reinterpret_cast<uint16_t*>(out), taps, cin, cout, cin_pad, cout_pad);
    LF_RETURN_LAUNCH();
}

extern "C" int lf_conv_bwd_data_fused(const lf_conv_desc* desc, const float* gy, const float* y_fwd,
                                      const float* rnorm_fwd, int fwd_act, float fwd_slope, int fwd_norm,
                                      const float* w_tc_packed, float* gx, void* stream) {
    if (desc == nullptr || desc->precision == 0 || !conv_tc_supported(desc)) {
        set_error("conv_bwd_data_fused: needs a tcgen05-supported shape and precision 1|2");
        return LF_EUNSUPPORTED;
    }
    const int q4 = ((desc->cin + 15) / 16 * 16) / 4;
    if (desc->cin % 16 != 0 || (q4 & (q4 - 1)) != 0 || q4 > 32) {
        set_error("conv_bwd_data_fused: Cin must be 16, 32, 64 or 128");
        return LF_EUNSUPPORTED;
    }
    LF_CHECK_ARG(gy && y_fwd && w_tc_packed && gx && (!fwd_norm || rnorm_fwd), "conv_bwd_data_fused: null pointer");
    TcPrologue pro{y_fwd, rnorm_fwd, fwd_act, fwd_norm, fwd_slope};
    return conv_tc_launch_ex(desc, gy, w_tc_packed, nullptr, gx, nullptr, &pro, nullptr, (cudaStream_t)stream);
}

namespace lf {
int expand_epi_supported(const lf_conv_desc* d);
int expand_epi_launch(const lf_conv_desc* d, const float* x, const float* w, float* y, const float* epi_y,
                      const float* epi_r, int epi_act, float epi_slope, int epi_norm, cudaStream_t st);
}

// 1 if lf_conv_bwd_data_epi can run this bwd-data descriptor with the fused backward epilogue
extern "C" int lf_conv_bwd_data_epi_supported(const lf_conv_desc* desc) {
    if (desc == nullptr) return 0;
    if (desc->ndim == -1) return lf::expand_epi_supported(desc);
    if (desc->precision == 0 || desc->act || desc->norm) return 0;
    tc::Plan pl;
    return (tc::make_plan(desc, pl) && (desc->cout & 3) == 0 && pl.cout_pad <= 32) ? 1 : 0;
}

// bwd-data convolution whose result row is pushed through the PixelNorm/LeakyReLU backward of the layer that
// produced the forward input (y_prev / rnorm_prev are that layer's saved output and norms): writes du_prev.
// `w` is the tcgen05-packed weight for conv descriptors (precision 1|2) and the fp32 [D][Cin][Cout] pack for the
// depth-expand (ndim -1).
extern "C" int lf_conv_bwd_data_epi(const lf_conv_desc* desc, const float* du, const float* w, const float* y_prev,
                                    const float* rnorm_prev, int prev_act, float prev_slope, int prev_norm,
                                    float* du_prev, void* stream) {
    if (!lf_conv_bwd_data_epi_supported(desc)) {
        set_error("conv_bwd_data_epi: unsupported shape/precision");
        return LF_EUNSUPPORTED;
    }
    LF_CHECK_ARG(du && w && y_prev && du_prev && (!prev_norm || rnorm_prev), "conv_bwd_data_epi: null pointer");
    if (desc->ndim == -1)
        return lf::expand_epi_launch(desc, du, w, du_prev, y_prev, rnorm_prev, prev_act, prev_slope, prev_norm,
                                     (cudaStream_t)stream);
    TcEpilogue epi{y_prev, rnorm_prev, prev_act, prev_norm, prev_slope};
    return conv_tc_launch_ex(desc, du, w, nullptr, du_prev, nullptr, nullptr, &epi, (cudaStream_t)stream);
}

extern "C" int lf_debug_tc_timeline(long long* host_out /* [3][64][2] */) {
    return (int)cudaMemcpyFromSymbol(host_out, tc::g_dbg, sizeof(long long) * 3 * 64 * 2);
}

// kernel launches lf_conv_fwd makes for this descriptor on the tcgen05 path: 1 (bf16), 2 (bf16x3 with the dual-weight
// first pass) or 3 (bf16x3 as three single-product passes); 0 if the shape is not covered
extern "C" int lf_conv_tc_passes(const lf_conv_desc* desc) {
    tc::Plan pl;
    if (desc == nullptr || desc->precision == 0 || !tc::make_plan(desc, pl)) return 0;
    if (desc->precision == 2) return 1;
    if (option(OPT_TC_NO_DUAL) == 0 && (desc->cout & 3) == 0 && pl.cout_pad <= 64) {
        tc::Plan pd;
        if (option(OPT_TC_NO_TRI) == 0 && desc->ndim == 2 && tc::make_plan(desc, pd, true, true)) return 1;
        if (tc::make_plan(desc, pd, true)) return 2;
    }
    return 3;
}

extern "C" int lf_conv_tc_supported(const lf_conv_desc* desc) {
    if (desc == nullptr) return 0;
    return conv_tc_supported(desc);
}
