// Depth-collapse projection (FactorProjection3d2d, modules/geometry.py:704-749; recon/models.py:436-452) on the tensor
// cores (sm_100a):   y[n, h, w, co] = PixelNorm(LeakyReLU(he * sum_{d, ci} x[n, d, h, w, ci] * W[d][ci][co] + b[co]))
//
// A GEMM with M = N*H*W positions, K = D*Cin (2048 at config B), N = Cout (32): 4.3 GF over 268 MB of input — HBM-bound
// (41 us) on tensor cores, FFMA-bound (126 us) on the exact kernel.  When the producing camera block left the
// split-planar twin of x (it does in the pose loop, for the fused backward of this layer: expand_tc.cu), one depth plane
// of two 128-position M-tiles is a 32 KB slab of 8 bulk TMA copies in the K-major canonical form, the plane's weight tile
// ([hi|lo][Cin/8][Cout][8], 4 KB) rides on the same barrier, and 6 MMAs per plane and tile (2 k-steps x the three bf16x3
// products) accumulate the whole depth axis into ONE [128 x Cout] TMEM tile per M-tile.  4-stage ring, one issuer warp
// per M-tile, epilogue as in the convolutions (scale, bias, LeakyReLU, PixelNorm thread-locally).
#include "tc_common.cuh"

#include <cuda_bf16.h>

namespace lf {
namespace ct {

using namespace tcx;

constexpr int kThreads = 384;      // warp 0: producer, 1..2: MMA issuers, 3: TMEM allocator, 4..11: epilogue (4 per M-tile)
constexpr int kStages = 4;

struct Params {
    const uint16_t* x;             // split-planar volume [hi|lo][n][d][KCi][PP][8]
    const uint16_t* wpk;           // [d][part][KCi][cout_pad rows][8] bf16
    const float* bias;
    float* y;                      // fp32 channels-last [n][h][w][cout]
    float* rnorm;                  // [n*h*w] (nullable)
    int64_t x_part;
    int n, d, h, w, Wp, PP, KCi, cout, cout_pad, T, NC, nprod;
    uint32_t a_bytes, btile_bytes, stage_bytes;
    float scale, slope;
    int act, norm;
    uint64_t magic_Wp;
};

__global__ void __launch_bounds__(kThreads, 1)
collapse_tc_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t s0 = smem_u32(smem);
    uint8_t* tail = smem + (size_t)kStages * p.stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);     // full[4] empty[4] acc_full[2] acc_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
    float* bias_s = reinterpret_cast<float*>(tmem_slot + 2);
    const uint32_t bar_f = smem_u32(bars), bar_e = bar_f + 8 * kStages;
    const uint32_t bar_accf = bar_e + 8 * kStages, bar_acce = bar_accf + 16;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    for (uint32_t i = threadIdx.x * 16; i < kStages * p.stage_bytes; i += kThreads * 16)      // rows past the plane stay finite
        *reinterpret_cast<uint4*>(smem + i) = make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x; i < p.cout_pad; i += kThreads) bias_s[i] = (p.bias != nullptr && i < p.cout) ? p.bias[i] : 0.f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(bar_f + 8 * i, 1); mbar_init(bar_e + 8 * i, 2); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_accf + 8 * i, 1); mbar_init(bar_acce + 8 * i, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int nparts = p.nprod == 3 ? 2 : 1;
    const uint32_t a_region = 256u * 16u;
    const int first = p.Wp + 1;
    const int items = p.n * p.NC;

    if (warp == 0) {
        // =========================== TMA PRODUCER: lanes 0..7 the slab regions, lane 8 the plane's weight tile ===========================
        uint32_t sc = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int col = item % p.NC, n = item / p.NC;
            const int start = first + col * 256;
            const uint32_t bytes = (uint32_t)min(256, p.PP - start) * 16u;
            for (int dd = 0; dd < p.d; ++dd, ++sc) {
                const uint32_t st = sc % kStages;
                if (lane == 0) {
                    mbar_wait(bar_e + 8 * st, ((sc / kStages) & 1) ^ 1, 41);
                    mbar_arrive_expect_tx(bar_f + 8 * st, bytes * (uint32_t)(p.KCi * nparts) + p.btile_bytes);
                }
                __syncwarp();
                if (lane < p.KCi * nparts) {
                    const int part = lane / p.KCi, kc = lane - part * p.KCi;
                    const uint16_t* src = p.x + part * p.x_part + ((((int64_t)n * p.d + dd) * p.KCi + kc) * p.PP + start) * 8;
                    bulk_g2s(s0 + st * p.stage_bytes + (uint32_t)lane * a_region, src, bytes, bar_f + 8 * st);
                } else if (lane == 8) {
                    bulk_g2s(s0 + st * p.stage_bytes + p.a_bytes, p.wpk + (int64_t)dd * (p.btile_bytes / 2), p.btile_bytes, bar_f + 8 * st);
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // =========================== MMA ISSUERS ===========================
        const int tile = warp - 1;
        const uint32_t idesc = idesc_bf16((uint32_t)p.cout_pad);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t a_lbo = ((a_region >> 4) << 16), b_lbo = (((uint32_t)p.cout_pad * 16u >> 4) << 16);
        const uint32_t a_part = (uint32_t)p.KCi * (a_region >> 4), b_part = ((uint32_t)p.KCi * p.cout_pad * 16u) >> 4;
        const int KS = p.KCi / 2;
        uint32_t sc = 0, it = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
            const int col = item % p.NC;
            const bool active = col * 2 + tile < p.T;
            mbar_wait(bar_acce + 8 * tile, (it & 1) ^ 1, 42);
            tc_fence_after();
            const uint32_t dcol = tmem_base + (uint32_t)tile * (uint32_t)p.cout_pad;
            for (int dd = 0; dd < p.d; ++dd, ++sc) {
                const uint32_t st = sc % kStages;
                mbar_wait(bar_f + 8 * st, (sc / kStages) & 1, 43);
                tc_fence_after();
                if (active) {
                    const uint32_t a0 = (a_lbo | ((s0 + st * p.stage_bytes) >> 4)) + (uint32_t)tile * 128u;
                    const uint32_t b0 = b_lbo | ((s0 + st * p.stage_bytes + p.a_bytes) >> 4);
                    for (int ks = 0; ks < KS; ++ks) {
                        const uint32_t ak = a0 + (uint32_t)ks * 2u * (a_region >> 4);
                        const uint32_t bk = b0 + (uint32_t)ks * 2u * ((uint32_t)p.cout_pad * 16u >> 4);
                        if (elect_one()) {
                            umma_f16(dcol, ak, desc_hi, bk, desc_hi, idesc, (dd == 0 && ks == 0) ? 0u : 1u);
                            if (p.nprod == 3) {
                                umma_f16(dcol, ak, desc_hi, bk + b_part, desc_hi, idesc, 1u);
                                umma_f16(dcol, ak + a_part, desc_hi, bk, desc_hi, idesc, 1u);
                            }
                        }
                    }
                }
                if (elect_one()) {
                    umma_commit(bar_e + 8 * st);
                    if (dd == p.d - 1) umma_commit(bar_accf + 8 * tile);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // =========================== EPILOGUE ===========================
        const int wq = warp & 3;
        const int tile = (warp - 4) >> 2;
        const float slope = p.act ? p.slope : 1.f;
        uint32_t it = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
            const int col = item % p.NC, n = item / p.NC;
            const int gt = col * 2 + tile;
            const int q = first + gt * 128 + wq * 32 + lane;
            const int yp = fast_div(q, p.magic_Wp), xp = q - yp * p.Wp;
            const bool valid = (gt < p.T) && (yp >= 1) && (yp <= p.h) && (xp >= 1) && (xp <= p.w);
            mbar_wait(bar_accf + 8 * tile, it & 1, 44);
            tc_fence_after();
            if (gt < p.T) {
                float v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(tile * p.cout_pad);
                tmem_ld16(taddr, v);
                if (p.cout_pad > 16) tmem_ld16(taddr + 16, v + 16);
                tmem_ld_wait();
                float ss = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float a = (i < p.cout_pad) ? v[i] * p.scale + bias_s[i] : 0.f;
                    a = fmaxf(a, a * slope);
                    if (i >= p.cout) a = 0.f;
                    v[i] = a;
                    ss += a * a;
                }
                if (valid) {
                    const int64_t pos = ((int64_t)n * p.h + (yp - 1)) * p.w + (xp - 1);
                    if (p.norm) {
                        const float rn = sqrtf(ss / (float)p.cout + 1e-8f), inv = 1.f / rn;
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] *= inv;
                        if (p.rnorm != nullptr) p.rnorm[pos] = rn;
                    }
                    float* yo = p.y + pos * p.cout;
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        if (i < p.cout) *reinterpret_cast<float4*>(yo + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
            tc_fence_before();
            mbar_arrive(bar_acce + 8 * tile);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 3) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64));
    }
}

// collapse weights [d][cin][cout] fp32 -> [d][part][Cin/8][cout_pad rows][8] bf16
__global__ void pack_weights_collapse_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int D, int cin, int cout,
                                             int KCi, int cout_pad) {
    const int64_t per_part = (int64_t)KCi * cout_pad * 8;
    const int64_t total = (int64_t)D * per_part;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = e;
        const int j = (int)(r % 8); r /= 8;
        const int row = (int)(r % cout_pad); r /= cout_pad;
        const int kc = (int)(r % KCi);
        const int dd = (int)(r / KCi);
        const int ci = kc * 8 + j;
        float v = 0.f;
        if (ci < cin && row < cout) v = w[((int64_t)dd * cin + ci) * cout + row];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const int64_t o = (int64_t)dd * 2 * per_part + ((int64_t)kc * cout_pad + row) * 8 + j;
        out[o] = __bfloat16_as_ushort(hi);
        out[o + per_part] = __bfloat16_as_ushort(lo);
    }
}

struct Plan {
    int cin_pad, cout_pad, KCi, Wp, PP, T, NC;
    uint32_t a_bytes, btile_bytes, stage_bytes, smem_bytes;
};

static bool make_plan(const lf_conv_desc* d, Plan& pl) {
    if (d->ndim != 1 || d->k != d->d) return false;
    if (d->precision != 1 && d->precision != 2) return false;
    if (d->n < 1 || d->d < 1 || d->h < 1 || d->w < 1 || d->cin < 1 || d->cout < 4 || (d->cout & 3)) return false;
    pl.cin_pad = (d->cin + 15) / 16 * 16;
    pl.cout_pad = (d->cout + 15) / 16 * 16;
    if (pl.cin_pad > 32 || pl.cout_pad > 32) return false;
    pl.KCi = pl.cin_pad / 8;
    pl.Wp = d->w + 2;
    pl.PP = (d->h + 2) * pl.Wp;
    if (pl.Wp >= 4096 || pl.PP >= (1 << 20)) return false;
    const int span = (d->h - 1) * pl.Wp + d->w;
    pl.T = (span + 127) / 128;
    pl.NC = (pl.T + 1) / 2;
    pl.a_bytes = 2u * pl.KCi * 256u * 16u;
    pl.btile_bytes = 2u * pl.KCi * pl.cout_pad * 16u;
    pl.stage_bytes = pl.a_bytes + pl.btile_bytes;
    pl.smem_bytes = kStages * pl.stage_bytes + 8 * (2 * kStages + 4) + 16 + 4 * 32 + 64;
    return pl.smem_bytes <= 227u * 1024u;
}

}  // namespace ct
}  // namespace lf

using namespace lf;

extern "C" int lf_collapse_tc_supported(const lf_conv_desc* desc) {
    ct::Plan pl;
    return (desc != nullptr && ct::make_plan(desc, pl)) ? 1 : 0;
}

extern "C" int64_t lf_collapse_tc_weight_bytes(int depth, int cin, int cout) {
    if (depth <= 0 || cin <= 0 || cout <= 0) return 0;
    const int64_t cin_pad = (cin + 15) / 16 * 16, cout_pad = (cout + 15) / 16 * 16;
    return (int64_t)depth * 2 * (cin_pad / 8) * cout_pad * 16;
}

extern "C" int lf_collapse_tc_pack_weights(const float* w /* [depth][cin][cout] */, void* out, int depth, int cin, int cout,
                                           void* stream) {
    LF_CHECK_ARG(w && out && depth > 0 && cin > 0 && cout > 0, "collapse_tc_pack_weights: bad arguments");
    const int cin_pad = (cin + 15) / 16 * 16, cout_pad = (cout + 15) / 16 * 16;
    const int64_t total = (int64_t)depth * (cin_pad / 8) * cout_pad * 8;
    ct::pack_weights_collapse_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        w, reinterpret_cast<uint16_t*>(out), depth, cin, cout, cin_pad / 8, cout_pad);
    LF_RETURN_LAUNCH();
}

extern "C" int lf_collapse_tc(const lf_conv_desc* desc, const void* x_split, const void* w_packed, const float* bias, float* y,
                              float* rnorm, void* stream) {
    ct::Plan pl;
    if (desc == nullptr || !ct::make_plan(desc, pl)) {
        set_error("collapse_tc: unsupported shape/precision (depth collapse with Cin <= 32, Cout <= 32, precision 1|2)");
        return LF_EUNSUPPORTED;
    }
    LF_CHECK_ARG(x_split && w_packed && y, "collapse_tc: null pointer");
    ct::Params p;
    p.x = reinterpret_cast<const uint16_t*>(x_split);
    p.wpk = reinterpret_cast<const uint16_t*>(w_packed);
    p.bias = bias; p.y = y; p.rnorm = rnorm;
    p.x_part = (int64_t)desc->n * desc->d * pl.cin_pad * pl.PP;
    p.n = desc->n; p.d = desc->d; p.h = desc->h; p.w = desc->w; p.Wp = pl.Wp; p.PP = pl.PP; p.KCi = pl.KCi;
    p.cout = desc->cout; p.cout_pad = pl.cout_pad; p.T = pl.T; p.NC = pl.NC; p.nprod = desc->precision == 1 ? 3 : 1;
    p.a_bytes = pl.a_bytes; p.btile_bytes = pl.btile_bytes; p.stage_bytes = pl.stage_bytes;
    p.scale = desc->scale; p.slope = desc->slope; p.act = desc->act; p.norm = desc->norm;
    p.magic_Wp = tcx::make_magic(pl.Wp);
    cudaError_t e = cudaFuncSetAttribute(ct::collapse_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (e != cudaSuccess) { set_error("collapse_tc: cannot raise dynamic smem: %s", cudaGetErrorString(e)); return (int)e; }
    const int items = desc->n * pl.NC;
    ct::collapse_tc_kernel<<<items < sm_count() ? items : sm_count(), ct::kThreads, pl.smem_bytes, (cudaStream_t)stream>>>(p);
    LF_RETURN_LAUNCH();
}
