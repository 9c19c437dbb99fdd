// Backward of the depth-collapse projection (FactorProjection3d2d, modules/geometry.py:704-749) w.r.t. its input volume,
// fused with the PixelNorm/LeakyReLU backward of the layer that PRODUCED that volume — on the tensor cores (sm_100a).
//
//   g[n, d, h, w, ci]  = he * sum_co du[n, h, w, co] * W[d][ci][co]                       (the "depth expand" GEMM)
//   du_prev[n, d, h, w, :] = gate(y) * (g - y * mean_c(g * y)) / r                       (lf_actnorm_bwd's formula)
//
// In the pose loop this pair ran as an FFMA expand kernel (0.17 ms, writes g) + lf_actnorm_bwd_split (0.15 ms, reads g
// and y, writes du_prev in split-planar form) at config B.  Here g never leaves the chip: per (sample, pair of
// 128-position M-tiles) the 2-D gradient du (K = Cout of the projection, 32) is staged once, the weights stream by
// groups of 8 depth slices as ONE N = 256 B tile ([W_d0 | ... | W_d7], 32 KB with its bf16 lo part), 6 MMAs per group
// (2 k-steps x the three bf16x3 products) fill a [128 x 256] TMEM tile, and the epilogue warps turn each 32-column slice
// into the split-planar du_prev of one depth plane (reading y from the producer's split-planar twin, 16 coalesced
// bytes per thread and 8-channel chunk).  HBM traffic = y twin in, du_prev twin out (2 x 286 MB at config B).
#include "tc_common.cuh"

#include <cuda_bf16.h>

namespace lf {
namespace ex {

using namespace tcx;

constexpr int kThreads = 384;      // warp 0: producer, 1..2: MMA issuers (one per M-tile), 3: TMEM allocator, 4..11: epilogue
constexpr int kBRing = 3;
constexpr int kDG = 8;             // depth slices per weight tile (N = kDG * cin_pad)

struct Params {
    const uint16_t* du;            // split-planar 2-D gradient [hi|lo][n][1][KCo][PP][8]
    const uint16_t* wpk;           // [depth group][part][KCo][kDG*cin_pad rows][8] bf16
    const uint16_t* yprev;         // producer layer's output, split-planar [hi|lo][n][d][KCi][PP][8]
    const float* rprev;            // its PixelNorm denominators [n*d*h*w] (nullable when !epi_norm)
    uint16_t* out;                 // du_prev, split-planar like yprev (halo zeros written)
    float* out32;                  // dense fp32 channels-last du_prev (nullable)
    int64_t du_part, y_part;
    int n, d, h, w, Wp, PP, KCo, KCi, cin, cin_pad, T, NC, ngroups, nprod;
    uint32_t a_bytes, btile_bytes;
    float scale, epi_slope;
    int epi_act, epi_norm;
    uint64_t magic_Wp;
};

__global__ void __launch_bounds__(kThreads, 1)
expand_tc_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t as0 = smem_u32(smem);
    const uint32_t bs0 = as0 + p.a_bytes;
    uint8_t* tail = smem + (size_t)p.a_bytes + (size_t)kBRing * p.btile_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);     // a_full, a_empty, b_full[3], b_empty[3], acc_full[2], acc_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 2 * kBRing + 4);
    const uint32_t bar_af = smem_u32(bars), bar_ae = bar_af + 8;
    const uint32_t bar_bf = bar_ae + 8, bar_be = bar_bf + 8 * kBRing;
    const uint32_t bar_accf = bar_be + 8 * kBRing, bar_acce = bar_accf + 16;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    for (uint32_t i = threadIdx.x * 16; i < p.a_bytes; i += kThreads * 16)       // rows past the plane stay finite
        *reinterpret_cast<uint4*>(smem + i) = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) {
        mbar_init(bar_af, 1); mbar_init(bar_ae, 2);
        for (int i = 0; i < kBRing; ++i) { mbar_init(bar_bf + 8 * i, 1); mbar_init(bar_be + 8 * i, 2); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_accf + 8 * i, 1); mbar_init(bar_acce + 8 * i, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int nparts = p.nprod == 3 ? 2 : 1;
    const uint32_t a_region = 256u * 16u;                           // one (part, k-chunk) of the staged du: 2 tiles x 128 rows
    const int first = p.Wp + 1;
    const int items = p.n * p.NC;

    if (warp == 0) {
        // =========================== TMA PRODUCER ===========================
        uint32_t ia = 0, sb = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++ia) {
            const int col = item % p.NC, n = item / p.NC;
            const int start = first + col * 256;
            const uint32_t bytes = (uint32_t)min(256, p.PP - start) * 16u;
            if (lane == 0) {
                mbar_wait(bar_ae, (ia & 1) ^ 1, 31);
                mbar_arrive_expect_tx(bar_af, bytes * (uint32_t)(p.KCo * nparts));
            }
            __syncwarp();
            if (lane < p.KCo * nparts) {
                const int part = lane / p.KCo, kc = lane - part * p.KCo;
                const uint16_t* src = p.du + part * p.du_part + (((int64_t)n * p.KCo + kc) * p.PP + start) * 8;
                bulk_g2s(as0 + (uint32_t)lane * a_region, src, bytes, bar_af);
            }
            for (int g = 0; g < p.ngroups; ++g, ++sb) {
                const uint32_t sl = sb % kBRing;
                if (lane == 8) {
                    mbar_wait(bar_be + 8 * sl, ((sb / kBRing) & 1) ^ 1, 32);
                    mbar_arrive_expect_tx(bar_bf + 8 * sl, p.btile_bytes);
                    bulk_g2s(bs0 + sl * p.btile_bytes, p.wpk + (int64_t)g * (p.btile_bytes / 2), p.btile_bytes, bar_bf + 8 * sl);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1 || warp == 2) {
        // =========================== MMA ISSUERS (one per M-tile) ===========================
        const int tile = warp - 1;
        const uint32_t NROWS = (uint32_t)(kDG * p.cin_pad);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t a_lbo = ((a_region >> 4) << 16), b_lbo = ((NROWS * 16u >> 4) << 16);
        const uint32_t a_part = (uint32_t)p.KCo * (a_region >> 4), b_part = ((uint32_t)p.KCo * NROWS * 16u) >> 4;
        const int KS = p.KCo / 2;
        uint32_t ia = 0, sb = 0, ig = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++ia) {
            const int col = item % p.NC;
            const bool active = col * 2 + tile < p.T;
            mbar_wait(bar_af, ia & 1, 33);
            const uint32_t a0 = (a_lbo | (as0 >> 4)) + (uint32_t)tile * 128u;
            for (int g = 0; g < p.ngroups; ++g, ++sb, ++ig) {
                const uint32_t sl = sb % kBRing;
                const int nd = min(kDG, p.d - g * kDG);
                mbar_wait(bar_acce + 8 * tile, (ig & 1) ^ 1, 34);            // the previous group's slices have been drained
                mbar_wait(bar_bf + 8 * sl, (sb / kBRing) & 1, 35);
                tc_fence_after();
                if (active) {
                    const uint32_t idesc = idesc_bf16((uint32_t)(nd * p.cin_pad));
                    const uint32_t b0 = b_lbo | ((bs0 + sl * p.btile_bytes) >> 4);
                    const uint32_t dcol = tmem_base + (uint32_t)tile * 256u;
                    for (int ks = 0; ks < KS; ++ks) {
                        const uint32_t ak = a0 + (uint32_t)ks * 2u * (a_region >> 4);
                        const uint32_t bk = b0 + (uint32_t)ks * 2u * (NROWS * 16u >> 4);
                        if (elect_one()) {
                            umma_f16(dcol, ak, desc_hi, bk, desc_hi, idesc, ks == 0 ? 0u : 1u);
                            if (p.nprod == 3) {
                                umma_f16(dcol, ak, desc_hi, bk + b_part, desc_hi, idesc, 1u);
                                umma_f16(dcol, ak + a_part, desc_hi, bk, desc_hi, idesc, 1u);
                            }
                        }
                    }
                }
                if (elect_one()) {
                    umma_commit(bar_be + 8 * sl);
                    umma_commit(bar_accf + 8 * tile);
                    if (g == p.ngroups - 1) umma_commit(bar_ae);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // =========================== EPILOGUE (group 0: tile 0, group 1: tile 1) ===========================
        const int wq = warp & 3;
        const int tile = (warp - 4) >> 2;
        uint32_t ig = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int col = item % p.NC, n = item / p.NC;
            const int gt = col * 2 + tile;
            const int q = first + gt * 128 + wq * 32 + lane;
            const int yp = fast_div(q, p.magic_Wp), xp = q - yp * p.Wp;
            const bool valid = (gt < p.T) && (yp >= 1) && (yp <= p.h) && (xp >= 1) && (xp <= p.w);
            uint4 yh4[4], yl4[4];                                  // the y row of the NEXT slice to process (prefetched)
            float rr = 1.f;
            for (int g = 0; g < p.ngroups; ++g, ++ig) {
                const int nd = min(kDG, p.d - g * kDG);
                // y rows are fetched one depth slice ahead of the slice being processed (packed hi / lo words): the epilogue
                // is otherwise one exposed HBM round trip per slice with only 256 threads per SM to hide it
                auto fetch = [&](int dd) {
                    const uint16_t* yh = p.yprev + ((((int64_t)n * p.d + dd) * p.KCi) * p.PP + q) * 8;
                    const uint16_t* yl = yh + p.y_part;
#pragma unroll
                    for (int kc = 0; kc < 4; ++kc) {
                        yh4[kc] = make_uint4(0u, 0u, 0u, 0u); yl4[kc] = yh4[kc];
                        if (kc < p.KCi) {
                            yh4[kc] = __ldg(reinterpret_cast<const uint4*>(yh + (int64_t)kc * p.PP * 8));
                            yl4[kc] = __ldg(reinterpret_cast<const uint4*>(yl + (int64_t)kc * p.PP * 8));
                        }
                    }
                    rr = p.epi_norm ? __ldg(p.rprev + (((int64_t)n * p.d + dd) * p.h + (yp - 1)) * p.w + (xp - 1)) : 1.f;
                };
                if (valid && g == 0) fetch(0);                 // (later groups: fetched by the previous slice)
                mbar_wait(bar_accf + 8 * tile, ig & 1, 36);
                tc_fence_after();
                for (int j = 0; j < nd; ++j) {
                    const int dd = g * kDG + j;
                    if (gt < p.T) {
                        float v[32];
                        const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(tile * 256 + j * p.cin_pad);
                        tmem_ld16(taddr, v);
                        if (p.cin_pad > 16) tmem_ld16(taddr + 16, v + 16);
                        tmem_ld_wait();
                        const int C = p.cin_pad;
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = (i < C) ? v[i] * p.scale : 0.f;
                        const int64_t rowbase = (((int64_t)n * p.d + dd) * p.KCi) * p.PP + q;
                        if (valid) {
                            // du_prev = gate(y) * (g - y * mean_c(g*y)) / r   (same formula as lf_actnorm_bwd)
                            float yv[32];
#pragma unroll
                            for (int kc = 0; kc < 4; ++kc) {
                                const uint32_t hw[4] = {yh4[kc].x, yh4[kc].y, yh4[kc].z, yh4[kc].w};
                                const uint32_t lw[4] = {yl4[kc].x, yl4[kc].y, yl4[kc].z, yl4[kc].w};
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    yv[kc * 8 + 2 * u] = __uint_as_float(hw[u] << 16) + __uint_as_float(lw[u] << 16);
                                    yv[kc * 8 + 2 * u + 1] = __uint_as_float(hw[u] & 0xffff0000u) + __uint_as_float(lw[u] & 0xffff0000u);
                                }
                            }
                            const float ir = 1.f / rr;
                            if (dd + 1 < p.d) fetch(dd + 1);            // next slice (possibly of the next weight group) in flight
                            float dot = 0.f;
                            if (p.epi_norm) {
#pragma unroll
                                for (int i = 0; i < 32; ++i) dot += v[i] * yv[i];
                                dot *= 1.f / (float)p.cin;
                            }
                            const float gs = p.epi_act ? p.epi_slope : 1.f;
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const float o = (v[i] - yv[i] * dot) * (p.epi_norm ? ir : 1.f);
                                v[i] = yv[i] > 0.f ? o : o * gs;
                            }
                            if (p.out32 != nullptr) {
                                float* yo = p.out32 + ((((int64_t)n * p.d + dd) * p.h + (yp - 1)) * p.w + (xp - 1)) * p.cin;
#pragma unroll
                                for (int i = 0; i < 32; i += 4)
                                    if (i < p.cin) *reinterpret_cast<float4*>(yo + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                            }
                        }
                        if (q < p.PP) {
                            uint16_t* hi = p.out + rowbase * 8;
                            uint16_t* lo = hi + p.y_part;
#pragma unroll
                            for (int kc = 0; kc < 4; ++kc) {
                                if (kc < p.KCi) {
                                    uint4 h4 = make_uint4(0u, 0u, 0u, 0u), l4 = h4;
                                    if (valid) {
                                        split_bf16x2(v[kc * 8 + 0], v[kc * 8 + 1], h4.x, l4.x);
                                        split_bf16x2(v[kc * 8 + 2], v[kc * 8 + 3], h4.y, l4.y);
                                        split_bf16x2(v[kc * 8 + 4], v[kc * 8 + 5], h4.z, l4.z);
                                        split_bf16x2(v[kc * 8 + 6], v[kc * 8 + 7], h4.w, l4.w);
                                    }
                                    *reinterpret_cast<uint4*>(hi + (int64_t)kc * p.PP * 8) = h4;
                                    *reinterpret_cast<uint4*>(lo + (int64_t)kc * p.PP * 8) = l4;
                                }
                            }
                        }
                    }
                    // halo positions no tile covers: [0, Wp+1) and [Wp+1 + T*128, PP)
                    if (tile == 0 && (col == 0 || col == p.NC - 1)) {
                        const int tid = wq * 32 + lane;
                        const int tail0 = first + p.T * 128;
                        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
                        for (int part = 0; part < 2; ++part)
                            for (int kc = 0; kc < p.KCi; ++kc) {
                                uint16_t* base = p.out + part * p.y_part + ((((int64_t)n * p.d + dd) * p.KCi + kc) * p.PP) * 8;
                                if (col == 0)
                                    for (int qq = tid; qq < first; qq += 128) *reinterpret_cast<uint4*>(base + (int64_t)qq * 8) = z4;
                                if (col == p.NC - 1)
                                    for (int qq = tail0 + tid; qq < p.PP; qq += 128) *reinterpret_cast<uint4*>(base + (int64_t)qq * 8) = z4;
                            }
                    }
                }
                tc_fence_before();
                mbar_arrive(bar_acce + 8 * tile);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 3) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// collapse weights [d][cin][cout] fp32 (lf_conv_fwd's packed layout for ndim 1) ->
// [depth group][part][k-chunk of 8 cout][row = (d % 8) * cin_pad + ci][8] bf16
__global__ void pack_weights_expand_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int D, int cin, int cout,
                                           int cin_pad, int KCo, int ngroups) {
    const int NROWS = kDG * cin_pad;
    const int64_t per_part = (int64_t)KCo * NROWS * 8;
    const int64_t total = (int64_t)ngroups * per_part;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = e;
        const int j = (int)(r % 8); r /= 8;
        const int row = (int)(r % NROWS); r /= NROWS;
        const int kc = (int)(r % KCo);
        const int g = (int)(r / KCo);
        const int dd = g * kDG + row / cin_pad, ci = row % cin_pad, co = kc * 8 + j;
        float v = 0.f;
        if (dd < D && ci < cin && co < cout) v = w[((int64_t)dd * cin + ci) * cout + co];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const int64_t o = (int64_t)g * 2 * per_part + ((int64_t)kc * NROWS + row) * 8 + j;
        out[o] = __bfloat16_as_ushort(hi);
        out[o + per_part] = __bfloat16_as_ushort(lo);
    }
}

struct Plan {
    int cin_pad, cout_pad, KCi, KCo, Wp, PP, T, NC, ngroups;
    uint32_t a_bytes, btile_bytes, smem_bytes;
};

// desc: the FORWARD collapse (ndim 1: cin = volume channels, cout = projected channels, d = k = depth)
static bool make_plan(const lf_conv_desc* d, Plan& pl) {
    if (d->ndim != 1 || d->k != d->d) return false;
    if (d->precision != 1 && d->precision != 2) return false;
    if (d->n < 1 || d->d < 1 || d->h < 1 || d->w < 1 || d->cin < 4 || (d->cin & 3) || d->cout < 1) return false;
    pl.cin_pad = (d->cin + 15) / 16 * 16;
    pl.cout_pad = (d->cout + 15) / 16 * 16;
    if (pl.cin_pad > 32 || pl.cout_pad > 64) return false;        // epilogue row in registers; du regions <= 16 lanes
    pl.KCi = pl.cin_pad / 8; pl.KCo = pl.cout_pad / 8;
    pl.Wp = d->w + 2;
    pl.PP = (d->h + 2) * pl.Wp;
    if (pl.Wp >= 4096 || pl.PP >= (1 << 20)) return false;
    const int span = (d->h - 1) * pl.Wp + d->w;
    pl.T = (span + 127) / 128;
    pl.NC = (pl.T + 1) / 2;
    pl.ngroups = (d->d + kDG - 1) / kDG;
    pl.a_bytes = 2u * pl.KCo * 256u * 16u;
    pl.btile_bytes = 2u * pl.KCo * (kDG * pl.cin_pad) * 16u;
    pl.smem_bytes = pl.a_bytes + kBRing * pl.btile_bytes + 8 * (2 + 2 * kBRing + 4) + 64;
    return pl.smem_bytes <= 227u * 1024u;
}

}  // namespace ex
}  // namespace lf

using namespace lf;

extern "C" int lf_expand_tc_supported(const lf_conv_desc* collapse_desc) {
    ex::Plan pl;
    return (collapse_desc != nullptr && ex::make_plan(collapse_desc, pl)) ? 1 : 0;
}

extern "C" int64_t lf_expand_tc_weight_bytes(int depth, int cin, int cout) {
    if (depth <= 0 || cin <= 0 || cout <= 0) return 0;
    const int64_t cin_pad = (cin + 15) / 16 * 16, cout_pad = (cout + 15) / 16 * 16;
    const int64_t groups = (depth + ex::kDG - 1) / ex::kDG;
    return groups * 2 * (cout_pad / 8) * (ex::kDG * cin_pad) * 16;
}

extern "C" int lf_expand_tc_pack_weights(const float* w /* [depth][cin][cout] */, void* out, int depth, int cin, int cout,
                                         void* stream) {
    LF_CHECK_ARG(w && out && depth > 0 && cin > 0 && cout > 0, "expand_tc_pack_weights: bad arguments");
    const int cin_pad = (cin + 15) / 16 * 16, cout_pad = (cout + 15) / 16 * 16;
    const int groups = (depth + ex::kDG - 1) / ex::kDG;
    const int64_t total = (int64_t)groups * (cout_pad / 8) * (ex::kDG * cin_pad) * 8;
    ex::pack_weights_expand_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        w, reinterpret_cast<uint16_t*>(out), depth, cin, cout, cin_pad, cout_pad / 8, groups);
    LF_RETURN_LAUNCH();
}

// du_split2d: split-planar twin of d(loss)/d(pre-activation of the collapse output) as a one-plane volume [n][1][h][w][cout];
// y_prev_split / rnorm_prev: the producer layer's output (split-planar [n][d][h][w][cin]) and norms;
// writes du_prev (split-planar, halo zeros included; dense fp32 too when du_prev32 is given).
extern "C" int lf_expand_tc_bwd_epi(const lf_conv_desc* collapse_desc, const void* du_split2d, const void* w_packed,
                                    const void* y_prev_split, const float* rnorm_prev, int prev_act, float prev_slope,
                                    int prev_norm, void* du_prev_split, float* du_prev32, void* stream) {
    ex::Plan pl;
    if (collapse_desc == nullptr || !ex::make_plan(collapse_desc, pl)) {
        set_error("expand_tc: unsupported shape/precision (depth collapse with Cin <= 32, Cout <= 64, precision 1|2)");
        return LF_EUNSUPPORTED;
    }
    LF_CHECK_ARG(du_split2d && w_packed && y_prev_split && du_prev_split && (!prev_norm || rnorm_prev), "expand_tc: null pointer");
    const lf_conv_desc* d = collapse_desc;
    ex::Params p;
    p.du = reinterpret_cast<const uint16_t*>(du_split2d);
    p.wpk = reinterpret_cast<const uint16_t*>(w_packed);
    p.yprev = reinterpret_cast<const uint16_t*>(y_prev_split);
    p.rprev = rnorm_prev;
    p.out = reinterpret_cast<uint16_t*>(du_prev_split);
    p.out32 = du_prev32;
    p.du_part = (int64_t)d->n * pl.cout_pad * pl.PP;
    p.y_part = (int64_t)d->n * d->d * pl.cin_pad * pl.PP;
    p.n = d->n; p.d = d->d; p.h = d->h; p.w = d->w; p.Wp = pl.Wp; p.PP = pl.PP; p.KCo = pl.KCo; p.KCi = pl.KCi;
    p.cin = d->cin; p.cin_pad = pl.cin_pad; p.T = pl.T; p.NC = pl.NC; p.ngroups = pl.ngroups;
    p.nprod = d->precision == 1 ? 3 : 1;
    p.a_bytes = pl.a_bytes; p.btile_bytes = pl.btile_bytes;
    p.scale = d->scale; p.epi_slope = prev_slope; p.epi_act = prev_act; p.epi_norm = prev_norm;
    p.magic_Wp = tcx::make_magic(pl.Wp);
    cudaError_t e = cudaFuncSetAttribute(ex::expand_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (e != cudaSuccess) { set_error("expand_tc: cannot raise dynamic smem: %s", cudaGetErrorString(e)); return (int)e; }
    const int items = d->n * pl.NC;
    ex::expand_tc_kernel<<<items < sm_count() ? items : sm_count(), ex::kThreads, pl.smem_bytes, (cudaStream_t)stream>>>(p);
    LF_RETURN_LAUNCH();
}
