// 3x3x3 Equalized convolution for WIDE layers on the 5th-gen tensor cores: weights streamed through shared memory (sm_100a).
//
// Reference op: modules/equalized.py:57-64 + blocks.py:152-158 at the released network widths
// (tools/train/train.sh:37-46: 64/128/256-channel camera and object blocks on a 16^3 latent).  The depth-batched
// kernel (conv3d_dz.cu) keeps all 27 taps' weights resident in shared memory, which stops at 32x32 channels
// (27*C^2 bf16 hi+lo = 7 MB at C = 256); here the weights stream:
//
//   item  = (sample n, output plane d, chunk c of NCH = 64|128 output channels)
//   D[tile t][128 positions, NCH] (+)= A[plane d+dz-1, 32-channel group g, tap (dy,dx)] * W[c][dz][g][tap]
//
// * A: one plane of one 32-channel group of the split-planar input ([hi|lo][n][d][C/8][(H+2)(W+2)][8 bf16]) is staged by
//   8 bulk TMA copies into the UMMA no-swizzle K-major form (as in conv3d_dz.cu); the 9 (dy,dx) taps of that slab are
//   descriptor start-address offsets; the whole (small) plane is covered by T <= 512/NCH M-tiles whose accumulators sit
//   side by side in TMEM and live across the whole K loop (3 dz x Cin/32 groups x 9 taps).
// * B: the weight tile of one (c, dz, g, tap) — [hi|lo][4 k-chunks][NCH rows][8] = 16 KB at NCH = 128 — is ONE bulk
//   copy out of the pre-packed buffer into a 4-slot ring; every tile is read by all T M-tiles (x3 bf16x3 products).
//   L2 -> SM weight traffic: 27*Cin*NCH*4 B per item (3.5 MB at 256 channels), ~14 B/clk/SM against 64-cycle MMAs.
// * 3 issuer warps (one M-tile each; a warp sustains one tcgen05.mma per ~120 cycles), 1 producer warp (one bulk copy
//   per lane), 4 epilogue warps: He scale + bias + LeakyReLU thread-locally, fp32 channels-last store, and the row's
//   partial sum of squares per channel chunk — PixelNorm needs the sum over ALL channels, so a small second kernel
//   (ws_finish_kernel) normalises the rows once every chunk has contributed.
#include "tc_common.cuh"

#include <cuda_bf16.h>

namespace lf {
namespace ws {

using namespace tcx;

constexpr int kThreads = 256;      // warp 0: producer, 1..3: MMA issuers (warp 2 also allocates TMEM), 4..7: epilogue
constexpr int kAStages = 2;
constexpr int kBRing = 4;
constexpr int kGroup = 32;         // input channels per streamed K block (4 k-chunks of 8)

struct Params {
    const uint16_t* x;             // split-planar input, hi part; lo part at + x_part
    const uint16_t* wpk;           // [chunk][dz][g][tap9][part][4][NCH][8] bf16
    const float* bias;
    float* y;                      // fp32 channels-last [n][d][h][w][cout]
    float* ss;                     // [positions][nchunks] partial sums of squares (nullable)
    int64_t x_part;
    int n, d, h, w, Wp, PP, KCin, G, cout, NCH, nchunks, T, L_alloc, items, nprod;
    int ndz, Tg, ngroups;          // 3 (3-D) | 1 (2-D: one plane per image); M-tiles per item and tile groups per plane
    uint32_t slab_bytes, btile_bytes;
    float scale, slope;
    int act;
    uint64_t magic_Wp;
};

__global__ void __launch_bounds__(kThreads, 1)
conv3d_ws_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t as0 = smem_u32(smem);
    const uint32_t bs0 = as0 + kAStages * p.slab_bytes;
    uint8_t* tail = smem + (size_t)kAStages * p.slab_bytes + (size_t)kBRing * p.btile_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);   // a_full[2] a_empty[2] b_full[4] b_empty[4] acc_full acc_empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kAStages + 2 * kBRing + 2);
    float* bias_s = reinterpret_cast<float*>(tmem_slot + 2);      // [NCH] of the current chunk (epilogue warps only)
    const uint32_t bar_af = smem_u32(bars), bar_ae = bar_af + 8 * kAStages;
    const uint32_t bar_bf = bar_ae + 8 * kAStages, bar_be = bar_bf + 8 * kBRing;
    const uint32_t bar_accf = bar_be + 8 * kBRing, bar_acce = bar_accf + 8;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    // rows of a slab past the end of the plane are never copied: they must hold finite values (their results are dropped)
    for (uint32_t i = threadIdx.x * 16; i < kAStages * p.slab_bytes; i += kThreads * 16)
        *reinterpret_cast<uint4*>(smem + i) = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) {
        for (int i = 0; i < kAStages; ++i) { mbar_init(bar_af + 8 * i, 1); mbar_init(bar_ae + 8 * i, 3); }
        for (int i = 0; i < kBRing; ++i) { mbar_init(bar_bf + 8 * i, 1); mbar_init(bar_be + 8 * i, 3); }
        mbar_init(bar_accf, 3);
        mbar_init(bar_acce, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t region = (uint32_t)p.L_alloc * 16u;              // one 8-channel chunk of a slab
    const int nparts = p.nprod == 3 ? 2 : 1;

    if (warp == 0) {
        // =========================== TMA PRODUCER ===========================
        // lanes 0..7: the 8 regions (part, k-chunk) of the activation slab; lane 8: the weight tile of each tap
        uint32_t sa = 0, sb = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            int r_ = item;
            const int c = r_ % p.nchunks; r_ /= p.nchunks;
            const int tg = r_ % p.ngroups; r_ /= p.ngroups;
            const int dpl = r_ % p.d, n = r_ / p.d;
            const int start = tg * p.Tg * 128;                               // first slab position of this tile group
            const uint32_t copy_bytes = (uint32_t)min(p.L_alloc, p.PP - start) * 16u;
            for (int dz = 0; dz < p.ndz; ++dz) {
                const int e = dpl + (p.ndz == 3 ? dz - 1 : 0);
                if (e < 0 || e >= p.d) continue;
                for (int g = 0; g < p.G; ++g, ++sa) {
                    const uint32_t st = sa % kAStages;
                    if (lane == 0) {
                        mbar_wait(bar_ae + 8 * st, ((sa / kAStages) & 1) ^ 1, 21);
                        mbar_arrive_expect_tx(bar_af + 8 * st, copy_bytes * 4u * (uint32_t)nparts);
                    }
                    __syncwarp();
                    if (lane < 4 * nparts) {
                        const int part = lane >> 2, kc = lane & 3;
                        const uint16_t* src = p.x + part * p.x_part +
                                              ((((int64_t)n * p.d + e) * p.KCin + g * 4 + kc) * p.PP + start) * 8;
                        bulk_g2s(as0 + st * p.slab_bytes + (uint32_t)lane * region, src, copy_bytes, bar_af + 8 * st);
                    }
                    const uint16_t* wt = p.wpk + ((((int64_t)c * p.ndz + dz) * p.G + g) * 9) * (p.btile_bytes / 2);
                    for (int tap = 0; tap < 9; ++tap, ++sb) {
                        const uint32_t sl = sb % kBRing;
                        if (lane == 8) {
                            mbar_wait(bar_be + 8 * sl, ((sb / kBRing) & 1) ^ 1, 22);
                            mbar_arrive_expect_tx(bar_bf + 8 * sl, p.btile_bytes);
                            bulk_g2s(bs0 + sl * p.btile_bytes, wt + (int64_t)tap * (p.btile_bytes / 2), p.btile_bytes, bar_bf + 8 * sl);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp <= 3) {
        // =========================== MMA ISSUERS: M-tiles iz, iz + 3 ===========================
        const int iz = warp - 1;
        const uint32_t idesc = idesc_bf16((uint32_t)p.NCH);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);                   // SBO = 128 B, version 1
        const uint32_t a_lbo = ((region >> 4) << 16), b_lbo = (((uint32_t)p.NCH * 16u >> 4) << 16);
        const uint32_t a_part = 4u * (region >> 4);                          // hi -> lo inside a slab
        const uint32_t b_part = (4u * (uint32_t)p.NCH * 16u) >> 4;           // hi -> lo inside a weight tile
        const uint32_t b_ks = (2u * (uint32_t)p.NCH * 16u) >> 4, a_ks = 2u * (region >> 4);   // +16 input channels
        uint32_t sa = 0, sb = 0, it = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
            const int tg = (item / p.nchunks) % p.ngroups;
            const int dpl = (item / (p.nchunks * p.ngroups)) % p.d;
            const int Tcur = min(p.Tg, p.T - tg * p.Tg);
            mbar_wait(bar_acce, (it & 1) ^ 1, 23);                           // the previous item's rows have been drained
            tc_fence_after();
            uint32_t fresh = 1;                                              // first MMA of the item overwrites
            for (int dz = 0; dz < p.ndz; ++dz) {
                const int e = dpl + (p.ndz == 3 ? dz - 1 : 0);
                if (e < 0 || e >= p.d) continue;
                for (int g = 0; g < p.G; ++g, ++sa) {
                    const uint32_t st = sa % kAStages;
                    mbar_wait(bar_af + 8 * st, (sa / kAStages) & 1, 24);
                    const uint32_t a0 = a_lbo | ((as0 + st * p.slab_bytes) >> 4);
                    for (int tap = 0; tap < 9; ++tap, ++sb) {
                        const uint32_t sl = sb % kBRing;
                        mbar_wait(bar_bf + 8 * sl, (sb / kBRing) & 1, 25);
                        tc_fence_after();
                        const int dy = tap / 3, dx = tap - dy * 3;
                        const uint32_t b0 = b_lbo | ((bs0 + sl * p.btile_bytes) >> 4);
                        for (int t = iz; t < Tcur; t += 3) {
                            const uint32_t dcol = tmem_base + (uint32_t)(t * p.NCH);
                            const uint32_t at = a0 + (uint32_t)(t * 128 + dy * p.Wp + dx);
                            for (int ks = 0; ks < 2; ++ks) {
                                const uint32_t ak = at + ks * a_ks, bk = b0 + ks * b_ks;
                                if (elect_one()) {
                                    umma_f16(dcol, ak, desc_hi, bk, desc_hi, idesc, (fresh && ks == 0) ? 0u : 1u);
                                    if (p.nprod == 3) {
                                        umma_f16(dcol, ak, desc_hi, bk + b_part, desc_hi, idesc, 1u);
                                        umma_f16(dcol, ak + a_part, desc_hi, bk, desc_hi, idesc, 1u);
                                    }
                                }
                            }
                        }
                        fresh = 0;
                        if (elect_one()) umma_commit(bar_be + 8 * sl);
                        __syncwarp();
                    }
                    if (elect_one()) umma_commit(bar_ae + 8 * st);
                    __syncwarp();
                }
            }
            if (elect_one()) umma_commit(bar_accf);
            __syncwarp();
        }
    } else if (warp >= 4) {
        // =========================== EPILOGUE ===========================
        const int wq = warp & 3;
        const float slope = p.act ? p.slope : 1.f;
        const int first = p.Wp + 1;
        uint32_t it = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
            int r_ = item;
            const int c = r_ % p.nchunks; r_ /= p.nchunks;
            const int tg = r_ % p.ngroups; r_ /= p.ngroups;
            const int dpl = r_ % p.d, n = r_ / p.d;
            const int Tcur = min(p.Tg, p.T - tg * p.Tg);
            for (int i = threadIdx.x - 128; i < p.NCH; i += 128) {
                const int ch = c * p.NCH + i;
                bias_s[i] = (p.bias != nullptr && ch < p.cout) ? p.bias[ch] : 0.f;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait(bar_accf, it & 1, 26);
            tc_fence_after();
            for (int t = 0; t < Tcur; ++t) {
                const int q = first + (tg * p.Tg + t) * 128 + wq * 32 + lane;
                const int yp = fast_div(q, p.magic_Wp), xp = q - yp * p.Wp;
                const bool valid = (yp >= 1) && (yp <= p.h) && (xp >= 1) && (xp <= p.w);
                const int64_t pos = (((int64_t)n * p.d + dpl) * p.h + (yp - 1)) * p.w + (xp - 1);
                float ssq = 0.f;
                for (int c0 = 0; c0 < p.NCH; c0 += 16) {
                    float v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(t * p.NCH + c0), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float a = v[i] * p.scale + bias_s[c0 + i];
                        a = fmaxf(a, a * slope);
                        v[i] = a;
                        ssq += a * a;
                    }
                    const int ch0 = c * p.NCH + c0;
                    if (valid) {
                        float* yo = p.y + pos * p.cout + ch0;
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            if (ch0 + i < p.cout) *reinterpret_cast<float4*>(yo + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    }
                }
                if (valid && p.ss != nullptr) p.ss[pos * p.nchunks + c] = ssq;
            }
            tc_fence_before();
            mbar_arrive(bar_acce);
            asm volatile("bar.sync 1, 128;" ::: "memory");           // bias_s is rewritten for the next item
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// PixelNorm over ALL channels once every chunk has written its partial sum of squares:
// y[pos][:] /= sqrt(sum_c ss[pos][c] / C + 1e-8); rnorm[pos] = that root (saved for the backward)
__global__ void __launch_bounds__(256)
ws_finish_kernel(float* __restrict__ y, const float* __restrict__ ss, float* __restrict__ rnorm, int64_t positions, int C,
                 int nchunks) {
    const int q4 = C >> 2;
    const int64_t units = positions * q4;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pos = u / q4;
        float s = 0.f;
        for (int c = 0; c < nchunks; ++c) s += __ldg(ss + pos * nchunks + c);
        const float rn = sqrtf(s / (float)C + 1e-8f), inv = 1.f / rn;
        float4* yp = reinterpret_cast<float4*>(y) + u;
        float4 v = *yp;
        v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
        *yp = v;
        if (rnorm != nullptr && u - pos * q4 == 0) rnorm[pos] = rn;
    }
}

// [27][cin][cout] fp32 -> [chunk][dz][g][tap9][part][kc4][NCH][8] bf16 (hi | lo)
__global__ void pack_weights_ws_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int cin, int cout, int G,
                                       int NCH, int nchunks, int ndz) {
    const int64_t per_tile = (int64_t)4 * NCH * 8;                 // elements of one part
    const int64_t total = (int64_t)nchunks * ndz * G * 9 * per_tile;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = e;
        const int j = (int)(r % 8); r /= 8;
        const int row = (int)(r % NCH); r /= NCH;
        const int kc = (int)(r % 4); r /= 4;
        const int tap = (int)(r % 9); r /= 9;
        const int g = (int)(r % G); r /= G;
        const int dz = (int)(r % ndz);
        const int c = (int)(r / ndz);
        const int ci = g * kGroup + kc * 8 + j, co = c * NCH + row;
        float v = 0.f;
        if (ci < cin && co < cout) v = w[((int64_t)(dz * 9 + tap) * cin + ci) * cout + co];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const int64_t tile = ((((int64_t)c * ndz + dz) * G + g) * 9 + tap) * (2 * per_tile);
        const int64_t o = tile + ((int64_t)kc * NCH + row) * 8 + j;
        out[o] = __bfloat16_as_ushort(hi);
        out[o + per_tile] = __bfloat16_as_ushort(lo);
    }
}

struct Plan {
    int cin_pad, cout_pad, G, NCH, nchunks, Wp, PP, T, L_alloc, ndz, Tg, ngroups;
    uint32_t slab_bytes, btile_bytes, smem_bytes;
};

static bool make_plan(const lf_conv_desc* d, Plan& pl) {
    if (d->k != 3 || !(d->ndim == 3 || (d->ndim == 2 && d->d == 1))) return false;
    pl.ndz = d->ndim == 3 ? 3 : 1;
    if (d->precision != 1 && d->precision != 2) return false;
    if (d->n < 1 || d->d < 1 || d->h < 1 || d->w < 1 || d->cin < 1 || d->cout < 4 || (d->cout & 3)) return false;
    pl.cin_pad = (d->cin + kGroup - 1) / kGroup * kGroup;
    if (((d->cin + 15) / 16 * 16) % kGroup != 0) return false;     // the split-planar input pads channels to 16: whole 32-groups only
    pl.cout_pad = (d->cout + 63) / 64 * 64;
    pl.G = pl.cin_pad / kGroup;
    pl.NCH = (pl.cout_pad % 128 == 0) ? 128 : 64;
    pl.nchunks = pl.cout_pad / pl.NCH;
    pl.Wp = d->w + 2;
    pl.PP = (d->h + 2) * pl.Wp;
    if (pl.Wp >= 4096 || pl.PP >= (1 << 20)) return false;
    const int span = (d->h - 1) * pl.Wp + d->w;
    pl.T = (span + 127) / 128;
    pl.btile_bytes = 2u * 4u * pl.NCH * 16u;
    // M-tiles per item: as many as TMEM (512 columns) and shared memory (two slabs of Tg*128 + 2 rows of halo) allow;
    // a larger plane is cut into tile groups, each of which streams the weights again
    const uint32_t fixed = kBRing * pl.btile_bytes + 8 * (2 * kAStages + 2 * kBRing + 2) + 16 + 4 * 128 + 64;
    pl.Tg = 0;
    for (int tg = (512 / pl.NCH < pl.T ? 512 / pl.NCH : pl.T); tg >= 1; --tg) {
        const int L = (tg * 128 + 2 * pl.Wp + 2 + 7) / 8 * 8;
        if ((uint32_t)L * 16u >= (1u << 18)) continue;              // descriptor LBO field
        const uint32_t slab = 8u * L * 16u;
        if (kAStages * slab + fixed > 227u * 1024u) continue;
        if (((kAStages * slab + kBRing * pl.btile_bytes) >> 4) >= (1u << 14)) continue;    // 14-bit start address
        pl.Tg = tg; pl.L_alloc = L; pl.slab_bytes = slab; pl.smem_bytes = kAStages * slab + fixed;
        break;
    }
    if (pl.Tg == 0) return false;
    pl.ngroups = (pl.T + pl.Tg - 1) / pl.Tg;
    if ((int64_t)d->n * d->d * pl.ngroups * pl.nchunks >= (1ll << 30)) return false;
    return true;
}

}  // namespace ws
}  // namespace lf

using namespace lf;

// wide 3x3x3 layers (Cin or Cout above what the depth-batched kernel keeps resident) on small planes (T * NCH <= 512)
extern "C" int lf_conv3d_ws_supported(const lf_conv_desc* desc) {
    ws::Plan pl;
    return (desc != nullptr && ws::make_plan(desc, pl)) ? 1 : 0;
}

extern "C" int64_t lf_conv3d_ws_weight_bytes(int taps, int cin, int cout) {
    if (cin <= 0 || cout <= 0 || (taps != 27 && taps != 9)) return 0;
    const int64_t cin_pad = (cin + 31) / 32 * 32, cout_pad = (cout + 63) / 64 * 64;
    return taps * cin_pad * cout_pad * 2 * 2;
}

// w: [27 | 9][Cin][Cout] fp32 (the packed-tap layout of lf_conv_fwd) -> the streamed tile order, bf16 hi | lo
extern "C" int lf_conv3d_ws_pack_weights(const float* w, void* out, int taps, int cin, int cout, void* stream) {
    LF_CHECK_ARG(w && out && cin > 0 && cout > 0 && (taps == 27 || taps == 9), "conv3d_ws_pack_weights: bad arguments");
    const int cin_pad = (cin + 31) / 32 * 32, cout_pad = (cout + 63) / 64 * 64;
    const int NCH = (cout_pad % 128 == 0) ? 128 : 64;
    const int64_t total = (int64_t)taps * cin_pad * cout_pad;
    ws::pack_weights_ws_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        w, reinterpret_cast<uint16_t*>(out), cin, cout, cin_pad / 32, NCH, cout_pad / NCH, taps / 9);
    LF_RETURN_LAUNCH();
}

// scratch floats for the PixelNorm partial sums
extern "C" int64_t lf_conv3d_ws_scratch(const lf_conv_desc* desc) {
    ws::Plan pl;
    if (desc == nullptr || !ws::make_plan(desc, pl) || !desc->norm) return 0;
    return (int64_t)desc->n * desc->d * desc->h * desc->w * pl.nchunks;
}

extern "C" int lf_conv3d_ws(const lf_conv_desc* desc, const void* x_split, const void* w_packed, const float* bias,
                            float* y32, float* rnorm, float* scratch, void* stream) {
    ws::Plan pl;
    if (desc == nullptr || !ws::make_plan(desc, pl)) {
        set_error("conv3d_ws: unsupported shape/precision (k=3, Cin padding to a multiple of 32, plane pitch that fits a slab)");
        return LF_EUNSUPPORTED;
    }
    LF_CHECK_ARG(x_split && w_packed && y32, "conv3d_ws: null pointer");
    LF_CHECK_ARG(!desc->norm || scratch, "conv3d_ws: PixelNorm needs the scratch buffer");
    cudaStream_t st = (cudaStream_t)stream;
    ws::Params p;
    p.x = reinterpret_cast<const uint16_t*>(x_split);
    p.wpk = reinterpret_cast<const uint16_t*>(w_packed);
    p.bias = bias; p.y = y32; p.ss = desc->norm ? scratch : nullptr;
    const int xin_pad = (desc->cin + 15) / 16 * 16;                 // the split-planar buffer's own channel padding
    p.x_part = (int64_t)desc->n * desc->d * xin_pad * pl.PP;
    p.n = desc->n; p.d = desc->d; p.h = desc->h; p.w = desc->w; p.Wp = pl.Wp; p.PP = pl.PP;
    p.KCin = xin_pad / 8; p.G = pl.G; p.cout = desc->cout; p.NCH = pl.NCH; p.nchunks = pl.nchunks; p.T = pl.T;
    p.L_alloc = pl.L_alloc; p.items = desc->n * desc->d * pl.ngroups * pl.nchunks; p.nprod = desc->precision == 1 ? 3 : 1;
    p.ndz = pl.ndz; p.Tg = pl.Tg; p.ngroups = pl.ngroups;
    p.slab_bytes = pl.slab_bytes; p.btile_bytes = pl.btile_bytes;
    p.scale = desc->scale; p.slope = desc->slope; p.act = desc->act;
    p.magic_Wp = tcx::make_magic(pl.Wp);
    cudaError_t e = cudaFuncSetAttribute(ws::conv3d_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (e != cudaSuccess) { set_error("conv3d_ws: cannot raise dynamic smem: %s", cudaGetErrorString(e)); return (int)e; }
    const int grid = p.items < sm_count() ? p.items : sm_count();
    ws::conv3d_ws_kernel<<<grid, ws::kThreads, pl.smem_bytes, st>>>(p);
    if (desc->norm) {
        const int64_t positions = (int64_t)desc->n * desc->d * desc->h * desc->w;
        const int64_t units = positions * (desc->cout >> 2);
        int64_t blocks = (units + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 16;
        if (blocks > cap) blocks = cap;
        ws::ws_finish_kernel<<<(unsigned)blocks, 256, 0, st>>>(y32, scratch, rnorm, positions, desc->cout, pl.nchunks);
    }
    LF_RETURN_LAUNCH();
}
