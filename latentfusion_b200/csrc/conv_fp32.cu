// Exact-fp32 implicit-GEMM convolution with the Equalized/Block epilogue fused (CUDA-core FFMA path).
//
// Replaces, per call, the reference's 7 memory passes per convolution
//   conv (cuDNN/MKLDNN) -> mul_(He const) -> add(bias) -> leaky_relu -> pow/mean/add/sqrt -> div
// (modules/equalized.py:57-64, modules/blocks.py:152-164, modules/__init__.py:14-15) by one kernel
// that reads x once and writes y (+ the per-position norm needed by backward) once.
//
// This is the bit-faithful fp32 path (precision = 0): used for the fp32-parity configuration, for odd
// channel counts, and as the in-library cross-check of the tcgen05 path (conv_tc.cu).
//
// GEMM view:  M = output positions, N = Cout, K = taps * Cin.
//   kind 0: regular 2-D / 3-D convolution, k in {1,3}, zero padding k/2, stride 1
//   kind 1: depth-collapse   x[N][D][H][W][Cin] -> y[N][H][W][Cout],  taps = D   (FactorProjection3d2d)
//   kind 2: depth-expand     x[N][H][W][Cin]    -> y[N][D][H][W][Cout], one 1x1 GEMM per depth slice,
//           bias is [D][Cout]                                                  (FactorProjection2d3d)
#include "common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace lf {

constexpr int BM = 128;   // positions per block
constexpr int BK = 16;    // K chunk (input channels per step)

struct ConvGeom {
    int kind;             // 0 regular, 1 collapse, 2 expand
    int n, d, h, w;       // INPUT extent for kind 0/1; for kind 2: d = output depth, input is [n][h][w]
    int cin, cout;
    int k, taps;          // taps: kind0 k^ndim, kind1 d, kind2 1
    int ndim;
    int64_t out_positions;
    float scale;
    int act; float slope; int norm;
};

// decode an output position into (n, z, y, x) for kind 0 (z = 0 for 2-D)
__device__ __forceinline__ void decode_pos(const ConvGeom& g, int64_t P, int& n, int& z, int& y, int& x) {
    x = (int)(P % g.w); P /= g.w;
    y = (int)(P % g.h); P /= g.h;
    if (g.kind == 0) { z = (int)(P % g.d); P /= g.d; }
    else if (g.kind == 2) { z = (int)(P % g.d); P /= g.d; }
    else z = 0;
    n = (int)P;
}

template <int BN, int AVEC>
__global__ void __launch_bounds__(128)
conv_fp32_kernel(const ConvGeom g, const float* __restrict__ x, const float* __restrict__ wp,
                 const float* __restrict__ bias, float* __restrict__ y, float* __restrict__ rnorm) {
    constexpr int TN = BN / 8;           // couts per thread
    __shared__ __align__(16) float sA[BK][BM + 4];
    __shared__ __align__(16) float sB[BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid & 7;              // cout group
    const int ty = tid >> 3;             // position group (8 positions each)
    const int n0 = blockIdx.y * BN;

    // this thread's A-load position
    const int64_t Pm = (int64_t)blockIdx.x * BM + tid;
    const bool pvalid = Pm < g.out_positions;
    int pn = 0, pz = 0, py = 0, px = 0;
    if (pvalid) decode_pos(g, Pm, pn, pz, py, px);
    const int pad = (g.kind == 0) ? g.k / 2 : 0;

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < g.taps; ++tap) {
        // input element offset (before channel) of this thread's position for this tap, or -1
        int64_t in_off = -1;
        const float* wt = wp;
        if (pvalid) {
            if (g.kind == 0) {
                int dz = 0, dy, dx;
                int t = tap;
                dx = t % g.k; t /= g.k;
                dy = t % g.k; t /= g.k;
                if (g.ndim == 3) dz = t; else dz = pad;
                const int iz = pz + dz - pad, iy = py + dy - pad, ix = px + dx - pad;
                if (iz >= 0 && iz < g.d && iy >= 0 && iy < g.h && ix >= 0 && ix < g.w)
                    in_off = ((((int64_t)pn * g.d + iz) * g.h + iy) * g.w + ix) * g.cin;
            } else if (g.kind == 1) {
                in_off = ((((int64_t)pn * g.d + tap) * g.h + py) * g.w + px) * g.cin;
            } else {
                in_off = (((int64_t)pn * g.h + py) * g.w + px) * g.cin;
            }
        }
        // weights: kind 2 selects the slice by output depth, which is NOT block-uniform -> handled below
        wt = wp + (int64_t)tap * g.cin * g.cout;

        for (int kc = 0; kc < g.cin; kc += BK) {
            // ---- A tile: sA[k][m] = x[pos(m) + tap][kc + k]
            if (AVEC == 4) {
#pragma unroll
                for (int q = 0; q < BK / 4; ++q) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (in_off >= 0 && kc + q * 4 < g.cin) v = ldg4(x + in_off + kc + q * 4);
                    sA[q * 4 + 0][tid] = v.x; sA[q * 4 + 1][tid] = v.y;
                    sA[q * 4 + 2][tid] = v.z; sA[q * 4 + 3][tid] = v.w;
                }
            } else {
#pragma unroll
                for (int q = 0; q < BK; ++q) {
                    float v = 0.f;
                    if (in_off >= 0 && kc + q < g.cin) v = __ldg(x + in_off + kc + q);
                    sA[q][tid] = v;
                }
            }
            // ---- B tile: sB[k][n] = w[tap][kc + k][n0 + n]
            for (int e = tid; e < BK * BN; e += 128) {
                const int kk = e / BN, nn = e - kk * BN;
                float v = 0.f;
                if (kc + kk < g.cin && n0 + nn < g.cout) v = __ldg(wt + (int64_t)(kc + kk) * g.cout + n0 + nn);
                sB[kk][nn] = v;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                float a[8], b[TN];
                const float4 a0 = *reinterpret_cast<const float4*>(&sA[kk][ty * 8]);
                const float4 a1 = *reinterpret_cast<const float4*>(&sA[kk][ty * 8 + 4]);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
                a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
                for (int j = 0; j < TN; j += 4) {
                    const float4 bv = *reinterpret_cast<const float4*>(&sB[kk][tx * TN + j]);
                    b[j] = bv.x; b[j + 1] = bv.y; b[j + 2] = bv.z; b[j + 3] = bv.w;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

    // ---- epilogue: scale, bias, LeakyReLU, PixelNorm (needs cout <= BN), store
    const float inv_c = 1.f / (float)g.cout;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t P = (int64_t)blockIdx.x * BM + ty * 8 + i;
        float v[TN];
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int co = n0 + tx * TN + j;
            float t = acc[i][j] * g.scale;
            if (bias != nullptr && co < g.cout) t += __ldg(bias + co);
            if (g.act) t = t > 0.f ? t : t * g.slope;
            if (co >= g.cout) t = 0.f;
            v[j] = t;
            ss += t * t;
        }
        if (g.norm) {
            ss += __shfl_xor_sync(0xffffffffu, ss, 1);
            ss += __shfl_xor_sync(0xffffffffu, ss, 2);
            ss += __shfl_xor_sync(0xffffffffu, ss, 4);
            const float r = sqrtf(ss * inv_c + 1e-8f);
#pragma unroll
            for (int j = 0; j < TN; ++j) v[j] = v[j] / r;
            if (rnorm != nullptr && tx == 0 && P < g.out_positions) rnorm[P] = r;
        }
        if (P < g.out_positions) {
            float* yp = y + P * g.cout + n0 + tx * TN;
            if ((g.cout & 3) == 0 && n0 + tx * TN + TN <= g.cout) {
#pragma unroll
                for (int j = 0; j < TN; j += 4)
                    *reinterpret_cast<float4*>(yp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j)
                    if (n0 + tx * TN + j < g.cout) yp[j] = v[j];
            }
        }
    }
}

// depth-expand (kind 2): y[n][t][h][w][co] = act(scale * sum_ci x[n][h][w][ci] * w[t][ci][co] + bias[t][co])
// One block = 128 (n,h,w) positions for ONE depth slice t (blockIdx.z) so the weight slice is uniform.
template <int BN, int AVEC>
__global__ void __launch_bounds__(128)
conv_expand_kernel(const ConvGeom g, const float* __restrict__ x, const float* __restrict__ wp,
                   const float* __restrict__ bias, float* __restrict__ y) {
    constexpr int TN = BN / 8;
    __shared__ __align__(16) float sA[BK][BM + 4];
    __shared__ __align__(16) float sB[BK][BN];
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
    const int n0 = blockIdx.y * BN;
    const int t = blockIdx.z;
    const int64_t hw = (int64_t)g.h * g.w;
    const int64_t in_positions = (int64_t)g.n * hw;
    const int64_t Pm = (int64_t)blockIdx.x * BM + tid;
    const int64_t in_off = Pm < in_positions ? Pm * g.cin : -1;
    const float* wt = wp + (int64_t)t * g.cin * g.cout;

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int kc = 0; kc < g.cin; kc += BK) {
        if (AVEC == 4) {
#pragma unroll
            for (int q = 0; q < BK / 4; ++q) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (in_off >= 0 && kc + q * 4 < g.cin) v = ldg4(x + in_off + kc + q * 4);
                sA[q * 4 + 0][tid] = v.x; sA[q * 4 + 1][tid] = v.y;
                sA[q * 4 + 2][tid] = v.z; sA[q * 4 + 3][tid] = v.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < BK; ++q) {
                float v = 0.f;
                if (in_off >= 0 && kc + q < g.cin) v = __ldg(x + in_off + kc + q);
                sA[q][tid] = v;
            }
        }
        for (int e = tid; e < BK * BN; e += 128) {
            const int kk = e / BN, nn = e - kk * BN;
            float v = 0.f;
            if (kc + kk < g.cin && n0 + nn < g.cout) v = __ldg(wt + (int64_t)(kc + kk) * g.cout + n0 + nn);
            sB[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[8], b[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&sA[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&sA[kk][ty * 8 + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                const float4 bv = *reinterpret_cast<const float4*>(&sB[kk][tx * TN + j]);
                b[j] = bv.x; b[j + 1] = bv.y; b[j + 2] = bv.z; b[j + 3] = bv.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t P = (int64_t)blockIdx.x * BM + ty * 8 + i;
        if (P >= in_positions) continue;
        const int64_t nb = P / hw, p2 = P - nb * hw;
        float* yp = y + (((nb * g.d + t) * hw) + p2) * g.cout;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int co = n0 + tx * TN + j;
            if (co >= g.cout) continue;
            float v = acc[i][j] * g.scale;
            if (bias != nullptr) v += __ldg(bias + (int64_t)t * g.cout + co);
            if (g.act) v = v > 0.f ? v : v * g.slope;
            yp[co] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The two depth projections of the render path (K5), as dedicated kernels.
//
// collapse (FactorProjection3d2d, geometry.py:744-749 + its conv): M = N*H*W positions, N = Cout <= 32,
// K = D*Cin (2048 at config B).  One 128-position tile per CTA gives only M/128 = 256 CTAs of 4 warps for
// 148 SMs, so the generic kernel above is latency-bound (12 TF/s).  Here a thread-block CLUSTER of KS CTAs
// shares one tile and splits the depth axis; the KS partial 128x32 accumulator tiles are summed through
// distributed shared memory in rank order (deterministic), each CTA finishing 128/KS rows of the tile
// including the fused scale/bias/LeakyReLU/PixelNorm epilogue.  Global->register prefetch of the next K
// chunk overlaps the FFMA loop (double-buffered shared tiles, one barrier per chunk).
// ---------------------------------------------------------------------------------------------
constexpr int RED_STRIDE = 36;     // floats per row of the partial-tile buffer (16-byte aligned rows)

constexpr int CK = 32;             // K chunk of the collapse kernel (input channels per stage)
constexpr int CMP = BM + 1;        // padded position pitch of its A tile (in float4 slots)

__global__ void __launch_bounds__(128, 4)
collapse_cluster_kernel(const ConvGeom g, const float* __restrict__ x, const float* __restrict__ wp,
                        const float* __restrict__ bias, float* __restrict__ y, float* __restrict__ rnorm, int KS) {
    // A tile, k-blocked: sA[stage][k/4][position] = 4 consecutive channels of one position (one 128-bit slot).
    // A quarter-warp stores the 8 slots (k/4 = 0..7) of ONE position: pitch 129 slots puts them in 8 different
    // 16-byte bank groups; the FFMA loop reads, per k/4, the slots of positions ty, ty+16, ... (4 adjacent slots
    // per warp, broadcast over the 8 cout lanes).  The partial-tile buffer of the cluster reduction aliases it.
    __shared__ __align__(16) float4 sA4[2][CK / 4][CMP];
    __shared__ __align__(16) float sB[2][CK][32];
    float* red = reinterpret_cast<float*>(&sA4[0][0][0]);
    static_assert(sizeof(float4) * 2 * (CK / 4) * CMP >= sizeof(float) * BM * RED_STRIDE, "partial tile must fit in the A buffer");
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();          // == blockIdx.z
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
    const int64_t hw = (int64_t)g.h * g.w;
    const int64_t tstride = hw * g.cin;
    const int tper = g.taps / KS, t0 = rank * tper;
    const int kchunks = (g.cin + CK - 1) / CK;
    const int nchunks = tper * kchunks;
    // global->shared mapping of the A tile: float4 slot e = q*128 + tid -> position e/8, channel quad e%8
    const int ak4 = tid & 7;
    uint32_t abase[BM / 16];          // in float4 units (host checks the tensor is < 2^32 of them); ~0u = no position
#pragma unroll
    for (int q = 0; q < BM / 16; ++q) {
        const int64_t P = (int64_t)blockIdx.x * BM + q * 16 + (tid >> 3);
        abase[q] = 0xffffffffu;
        if (P < g.out_positions) { const int64_t nb = P / hw; abase[q] = (uint32_t)((((nb * g.d) * hw + (P - nb * hw)) * g.cin) >> 2); }
    }
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const uint32_t tstride4 = (uint32_t)(tstride >> 2);

    float4 ra[BM / 16];
    float4 rb[CK * 8 / 128];
    auto load_chunk = [&](int c) {
        const int tap = t0 + c / kchunks, kc = (c % kchunks) * CK;
        const int k = kc + ak4 * 4;
#pragma unroll
        for (int q = 0; q < BM / 16; ++q)
            ra[q] = (abase[q] != 0xffffffffu && k < g.cin) ? __ldg(x4 + (abase[q] + (uint32_t)tap * tstride4 + (uint32_t)(k >> 2))) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < CK * 8 / 128; ++q) {
            const int e = tid + q * 128, kk = e >> 3, nn = (e & 7) * 4;
            const float* wt = wp + ((int64_t)tap * g.cin + kc + kk) * g.cout + nn;
            rb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kc + kk < g.cin) {
                if ((g.cout & 3) == 0 && nn + 3 < g.cout) rb[q] = ldg4(wt);
                else {
                    if (nn + 0 < g.cout) rb[q].x = __ldg(wt + 0);
                    if (nn + 1 < g.cout) rb[q].y = __ldg(wt + 1);
                    if (nn + 2 < g.cout) rb[q].z = __ldg(wt + 2);
                    if (nn + 3 < g.cout) rb[q].w = __ldg(wt + 3);
                }
            }
        }
    };
    auto store_chunk = [&](int st) {
#pragma unroll
        for (int q = 0; q < BM / 16; ++q) sA4[st][ak4][q * 16 + (tid >> 3)] = ra[q];
#pragma unroll
        for (int q = 0; q < CK * 8 / 128; ++q) {
            const int e = tid + q * 128;
            *reinterpret_cast<float4*>(&sB[st][e >> 3][(e & 7) * 4]) = rb[q];
        }
    };

    float acc[8][4];                 // rows ty + 16*i, couts tx*4 + j
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    load_chunk(0);
    store_chunk(0);
    __syncthreads();
    for (int c = 0; c < nchunks; ++c) {
        const int st = c & 1;
        if (c + 1 < nchunks) load_chunk(c + 1);
#pragma unroll
        for (int k4 = 0; k4 < CK / 4; ++k4) {
            float4 a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = sA4[st][k4][ty + 16 * i];
#pragma unroll
            for (int kq = 0; kq < 4; ++kq) {
                const float4 bv = *reinterpret_cast<const float4*>(&sB[st][k4 * 4 + kq][tx * 4]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float av = kq == 0 ? a[i].x : kq == 1 ? a[i].y : kq == 2 ? a[i].z : a[i].w;
                    ffma2_bcast(av, bv.x, bv.y, acc[i][0], acc[i][1]);
                    ffma2_bcast(av, bv.z, bv.w, acc[i][2], acc[i][3]);
                }
            }
        }
        if (c + 1 < nchunks) store_chunk(st ^ 1);
        __syncthreads();
    }

    // ---- cluster reduction through distributed shared memory (the A buffer is free now)
#pragma unroll
    for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(&red[(ty + 16 * i) * RED_STRIDE + tx * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    cluster.sync();
    const int rows = BM / KS;
    const float inv_c = 1.f / (float)g.cout;
    for (int rr = ty; rr < rows; rr += 16) {
        const int row = rank * rows + rr;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < KS; ++r) {                       // fixed rank order: bit-reproducible
            const float* peer = cluster.map_shared_rank(red, r);
            const float4 pv = *reinterpret_cast<const float4*>(peer + row * RED_STRIDE + tx * 4);
            v.x += pv.x; v.y += pv.y; v.z += pv.z; v.w += pv.w;
        }
        const int64_t P = (int64_t)blockIdx.x * BM + row;
        float o[4] = {v.x, v.y, v.z, v.w};
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = tx * 4 + j;
            float t = o[j] * g.scale;
            if (bias != nullptr && co < g.cout) t += __ldg(bias + co);
            if (g.act) t = t > 0.f ? t : t * g.slope;
            if (co >= g.cout) t = 0.f;
            o[j] = t;
            ss += t * t;
        }
        if (g.norm) {
            ss += __shfl_xor_sync(0xffffffffu, ss, 1);
            ss += __shfl_xor_sync(0xffffffffu, ss, 2);
            ss += __shfl_xor_sync(0xffffffffu, ss, 4);
            const float r = sqrtf(ss * inv_c + 1e-8f);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = o[j] / r;
            if (rnorm != nullptr && tx == 0 && P < g.out_positions) rnorm[P] = r;
        }
        if (P < g.out_positions) {
            float* yp = y + P * g.cout + tx * 4;
            if ((g.cout & 3) == 0 && tx * 4 + 4 <= g.cout) *reinterpret_cast<float4*>(yp) = make_float4(o[0], o[1], o[2], o[3]);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (tx * 4 + j < g.cout) yp[j] = o[j];
            }
        }
    }
    cluster.sync();          // nobody leaves while a peer may still read its partial tile
}

// expand (FactorProjection2d3d, geometry.py:724-728; and the backward of the collapse): per depth slice t a
// 1x1 GEMM with its own weight slice, y[n][t][p][co] = act(scale * sum_ci x[n][p][ci] * w[t][ci][co] + b[t][co]).
// The 128 x Cin input tile is staged in shared memory ONCE per CTA and reused for TS depth slices (the
// generic kernel re-staged it per slice); weight slices are register-prefetched; 128-bit output stores.
template <int CINMAX>
__global__ void __launch_bounds__(128)
expand_multi_kernel(const ConvGeom g, const float* __restrict__ x, const float* __restrict__ wp,
                    const float* __restrict__ bias, float* __restrict__ y, int TS,
                    const float* __restrict__ epi_y, const float* __restrict__ epi_r, int epi_act, int epi_norm,
                    float epi_slope) {
    __shared__ __align__(16) float sA[CINMAX][BM + 4];
    __shared__ __align__(16) float sB[2][CINMAX][32];
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
    const int n0 = blockIdx.y * 32;
    const int64_t hw = (int64_t)g.h * g.w;
    const int64_t in_positions = (int64_t)g.n * hw;
    const int64_t Pm = (int64_t)blockIdx.x * BM + tid;
    const int t0 = blockIdx.z * TS, t1 = min(g.d, t0 + TS);
    // stage the input tile: sA[ci][pos]
    for (int kc = 0; kc < g.cin; kc += 4) {
        const float4 v = (Pm < in_positions) ? ldg4(x + Pm * g.cin + kc) : make_float4(0.f, 0.f, 0.f, 0.f);
        sA[kc + 0][tid] = v.x; sA[kc + 1][tid] = v.y; sA[kc + 2][tid] = v.z; sA[kc + 3][tid] = v.w;
    }
    const int nb_elems = g.cin * 8;                       // float4 slots of one weight slice tile [cin][32]
    float4 rb[CINMAX * 8 / 128];
    auto load_b = [&](int t) {
#pragma unroll
        for (int q = 0; q < CINMAX * 8 / 128; ++q) {
            const int e = tid + q * 128;
            rb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < nb_elems) {
                const int kk = e >> 3, nn = (e & 7) * 4;
                const float* wt = wp + ((int64_t)t * g.cin + kk) * g.cout + n0 + nn;
                if ((g.cout & 3) == 0 && n0 + nn + 3 < g.cout) rb[q] = ldg4(wt);
                else {
                    if (n0 + nn + 0 < g.cout) rb[q].x = __ldg(wt + 0);
                    if (n0 + nn + 1 < g.cout) rb[q].y = __ldg(wt + 1);
                    if (n0 + nn + 2 < g.cout) rb[q].z = __ldg(wt + 2);
                    if (n0 + nn + 3 < g.cout) rb[q].w = __ldg(wt + 3);
                }
            }
        }
    };
    auto store_b = [&](int st) {
#pragma unroll
        for (int q = 0; q < CINMAX * 8 / 128; ++q) {
            const int e = tid + q * 128;
            if (e < nb_elems) *reinterpret_cast<float4*>(&sB[st][e >> 3][(e & 7) * 4]) = rb[q];
        }
    };
    load_b(t0);
    store_b(0);
    __syncthreads();
    for (int t = t0; t < t1; ++t) {
        const int st = (t - t0) & 1;
        if (t + 1 < t1) load_b(t + 1);
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int kk = 0; kk < g.cin; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&sA[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&sA[kk][ty * 8 + 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&sB[st][kk][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                ffma2_bcast(a[i], bv.x, bv.y, acc[i][0], acc[i][1]);
                ffma2_bcast(a[i], bv.z, bv.w, acc[i][2], acc[i][3]);
            }
        }
        float bj[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = n0 + tx * 4 + j;
            if (bias != nullptr && co < g.cout) bj[j] = __ldg(bias + (int64_t)t * g.cout + co);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t P = (int64_t)blockIdx.x * BM + ty * 8 + i;
            const bool live = P < in_positions;                 // (no early exit: the fused epilogue shuffles)
            const int64_t Pc = live ? P : in_positions - 1;
            const int64_t nb = Pc / hw, p2 = Pc - nb * hw;
            const int64_t orow = ((nb * g.d + t) * hw) + p2;      // output position index
            float* yp = y + orow * g.cout + n0 + tx * 4;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float u = acc[i][j] * g.scale + bj[j];
                if (g.act) u = u > 0.f ? u : u * g.slope;
                v[j] = u;
            }
            if (epi_y != nullptr) {
                // backward of the layer that produced this tensor's forward twin (grid.y == 1, Cout % 4 == 0 checked on
                // the host): du = gate(y) * (g - y * mean_c(g*y)) / r, the 8 lanes tx of a row hold its Cout channels
                const bool ch = tx * 4 < g.cout;
                const float4 y4 = ch ? ldg4(epi_y + orow * g.cout + tx * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
                float dot = 0.f, ir = 1.f;
                if (epi_norm) {
                    dot = ch ? (v[0] * yv[0] + v[1] * yv[1] + v[2] * yv[2] + v[3] * yv[3]) : 0.f;
                    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
                    dot *= 1.f / (float)g.cout;
                    ir = 1.f / __ldg(epi_r + orow);
                }
                const float gs = epi_act ? epi_slope : 1.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float o = (v[j] - yv[j] * dot) * ir;
                    v[j] = yv[j] > 0.f ? o : o * gs;
                }
            }
            if (!live) continue;
            if ((g.cout & 3) == 0 && n0 + tx * 4 + 4 <= g.cout) st4_stream(yp, make_float4(v[0], v[1], v[2], v[3]));
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n0 + tx * 4 + j < g.cout) yp[j] = v[j];
            }
        }
        if (t + 1 < t1) store_b(st ^ 1);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// PixelNorm over a (gd x C) group:  x[(o*gd + t)*inner + p][c], group = all (t, c) for fixed (o, p).
// gd = 1 is the ordinary per-position PixelNorm (modules/__init__.py:14-15); gd = S is the
// normalisation over the C*S channels of FactorProjection2d3d before its view() (geometry.py:724-728).
// One warp per group.
// ---------------------------------------------------------------------------------------------
__global__ void pixelnorm_group_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ rnorm,
                                       int64_t outer, int gd, int64_t inner, int c) {
    const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (grp >= outer * inner) return;
    const int64_t o = grp / inner, p = grp - o * inner;
    const int64_t len = (int64_t)gd * c;
    float ss = 0.f;
    for (int64_t e = lane; e < len; e += 32) {
        const int64_t t = e / c, cc = e - t * c;
        const float v = x[((o * gd + t) * inner + p) * c + cc];
        ss += v * v;
    }
    ss = warp_sum(ss);
    const float r = sqrtf(ss / (float)len + 1e-8f);
    for (int64_t e = lane; e < len; e += 32) {
        const int64_t t = e / c, cc = e - t * c;
        const int64_t idx = ((o * gd + t) * inner + p) * c + cc;
        y[idx] = x[idx] / r;
    }
    if (lane == 0 && rnorm != nullptr) rnorm[grp] = r;
}

// du = LeakyReLU'( . ) * PixelNorm^T(gy):  da = (gy - y * mean(gy*y)) / r ;  du = da * (y > 0 ? 1 : slope)
__global__ void actnorm_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ y,
                                   const float* __restrict__ rnorm, float* __restrict__ du,
                                   int64_t outer, int gd, int64_t inner, int c, int act, float slope, int norm) {
    const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (grp >= outer * inner) return;
    const int64_t o = grp / inner, p = grp - o * inner;
    const int64_t len = (int64_t)gd * c;
    float dot = 0.f, r = 1.f;
    if (norm) {
        for (int64_t e = lane; e < len; e += 32) {
            const int64_t t = e / c, cc = e - t * c;
            const int64_t idx = ((o * gd + t) * inner + p) * c + cc;
            dot += gy[idx] * y[idx];
        }
        dot = warp_sum(dot) / (float)len;
        r = rnorm[grp];
    }
    for (int64_t e = lane; e < len; e += 32) {
        const int64_t t = e / c, cc = e - t * c;
        const int64_t idx = ((o * gd + t) * inner + p) * c + cc;
        const float yv = y[idx];
        float g = gy[idx];
        if (norm) g = (g - yv * dot) / r;
        if (act) g = yv > 0.f ? g : g * slope;
        du[idx] = g;
    }
}

// fast path of the above for ordinary layers (gd == 1, C/4 a power of two <= 32): a position's channels sit on
// C/4 consecutive lanes as float4s, the channel mean is an xor-shuffle over that lane group, every access is a
// coalesced 128-bit load/store (the generic kernel above walks a position with one warp and scalar accesses).
__global__ void __launch_bounds__(256)
actnorm_bwd_vec_kernel(const float* __restrict__ gy, const float* __restrict__ y, const float* __restrict__ rnorm,
                       float* __restrict__ du, int64_t positions, int c, int lg, int act, float slope, int norm) {
    const int q4 = c >> 2;
    const int64_t units = positions * q4;
    const int64_t units_pad = (units + 31) & ~(int64_t)31;
    const float inv_c = 1.f / (float)c;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < units_pad; u += (int64_t)gridDim.x * blockDim.x) {
        const bool live = u < units;
        const int64_t uu = live ? u : units - 1;
        const float4 g = ldg4(gy + uu * 4), yv = ldg4(y + uu * 4);
        float4 o = g;
        if (norm) {
            float dot = g.x * yv.x + g.y * yv.y + g.z * yv.z + g.w * yv.w;
            for (int s = 1; s < (1 << lg); s <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, s);
            dot *= inv_c;
            const float ir = 1.f / __ldg(rnorm + (uu >> lg));
            o.x = (g.x - yv.x * dot) * ir; o.y = (g.y - yv.y * dot) * ir;
            o.z = (g.z - yv.z * dot) * ir; o.w = (g.w - yv.w * dot) * ir;
        }
        if (act) {
            o.x = yv.x > 0.f ? o.x : o.x * slope; o.y = yv.y > 0.f ? o.y : o.y * slope;
            o.z = yv.z > 0.f ? o.z : o.z * slope; o.w = yv.w > 0.f ? o.w : o.w * slope;
        }
        if (live) *reinterpret_cast<float4*>(du + u * 4) = o;
    }
}

// ---------------------------------------------------------------------------------------------
// weight / bias gradient (training).  grad_w[tap][ci][co] += scale * sum_pos x[pos+tap][ci] * du[pos][co]
// Each block takes a chunk of positions, loops over taps; partial [Cin x Cout] products reduced with
// fp32 atomics (order non-deterministic; training only).
// ---------------------------------------------------------------------------------------------
constexpr int WG_CHUNK = 64;   // positions per inner step

__global__ void __launch_bounds__(256)
conv_bwd_weight_kernel(const ConvGeom g, const float* __restrict__ x, const float* __restrict__ du,
                       float* __restrict__ gw, float* __restrict__ gb, int64_t pos_per_block) {
    extern __shared__ float sm[];
    float* sX = sm;                                // [WG_CHUNK][cin]
    float* sD = sm + (size_t)WG_CHUNK * g.cin;     // [WG_CHUNK][cout]
    const int tid = threadIdx.x;
    const int64_t p_begin = (int64_t)blockIdx.x * pos_per_block;
    const int64_t p_end = min(g.out_positions, p_begin + pos_per_block);
    const int pad = (g.kind == 0) ? g.k / 2 : 0;
    const int pairs = g.cin * g.cout;
    const int64_t hw = (int64_t)g.h * g.w;

    for (int tap = 0; tap < g.taps; ++tap) {
        int dz = 0, dy = 0, dx = 0;
        if (g.kind == 0) {
            int t = tap;
            dx = t % g.k; t /= g.k; dy = t % g.k; t /= g.k;
            dz = (g.ndim == 3) ? t : pad;
        }
        // register accumulators: each thread owns pairs tid, tid+256, ... (up to 16)
        float acc[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[q] = 0.f;
        float bacc = 0.f;
        for (int64_t p0 = p_begin; p0 < p_end; p0 += WG_CHUNK) {
            const int cnt = (int)min((int64_t)WG_CHUNK, p_end - p0);
            for (int e = tid; e < cnt * g.cin; e += 256) {
                const int m = e / g.cin, ci = e - m * g.cin;
                const int64_t P = p0 + m;
                int n, z, yy, xx;
                float v = 0.f;
                if (g.kind == 0) {
                    decode_pos(g, P, n, z, yy, xx);
                    const int iz = z + dz - pad, iy = yy + dy - pad, ix = xx + dx - pad;
                    if (iz >= 0 && iz < g.d && iy >= 0 && iy < g.h && ix >= 0 && ix < g.w)
                        v = x[((((int64_t)n * g.d + iz) * g.h + iy) * g.w + ix) * g.cin + ci];
                } else if (g.kind == 1) {
                    const int64_t nb = P / hw, p2 = P - nb * hw;
                    v = x[((nb * g.d + tap) * hw + p2) * g.cin + ci];
                } else {
                    v = x[P * g.cin + ci];          // P enumerates (n,h,w)
                }
                sX[m * g.cin + ci] = v;
            }
            for (int e = tid; e < cnt * g.cout; e += 256) {
                const int m = e / g.cout, co = e - m * g.cout;
                const int64_t P = p0 + m;
                float v;
                if (g.kind == 2) {
                    const int64_t nb = P / hw, p2 = P - nb * hw;
                    v = du[((nb * g.d + tap) * hw + p2) * g.cout + co];
                } else {
                    v = du[P * g.cout + co];
                }
                sD[m * g.cout + co] = v;
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int pr = tid + q * 256;
                if (pr < pairs) {
                    const int ci = pr / g.cout, co = pr - ci * g.cout;
                    float a = 0.f;
                    for (int m = 0; m < cnt; ++m) a = fmaf(sX[m * g.cin + ci], sD[m * g.cout + co], a);
                    acc[q] += a;
                }
            }
            if (gb != nullptr && tid < g.cout && (g.kind == 2 || tap == 0)) {
                float a = 0.f;
                for (int m = 0; m < cnt; ++m) a += sD[m * g.cout + tid];
                bacc += a;
            }
            __syncthreads();
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int pr = tid + q * 256;
            if (pr < pairs) atomicAdd(gw + (int64_t)tap * pairs + pr, acc[q] * g.scale);
        }
        if (gb != nullptr && tid < g.cout && (g.kind == 2 || tap == 0))
            atomicAdd(gb + (g.kind == 2 ? (int64_t)tap * g.cout : 0) + tid, bacc);
    }
}

static int make_geom(const lf_conv_desc* d, ConvGeom& g) {
    LF_CHECK_ARG(d != nullptr, "conv: null descriptor");
    LF_CHECK_ARG(d->n > 0 && d->h > 0 && d->w > 0 && d->d > 0 && d->cin > 0 && d->cout > 0, "conv: bad extents");
    g.n = d->n; g.d = d->d; g.h = d->h; g.w = d->w; g.cin = d->cin; g.cout = d->cout;
    g.k = d->k; g.scale = d->scale; g.act = d->act; g.slope = d->slope; g.norm = d->norm;
    g.ndim = d->ndim;
    if (d->ndim == 2 || d->ndim == 3) {
        LF_CHECK_ARG(d->k == 1 || d->k == 3, "conv: kernel size %d unsupported (1 or 3)", d->k);
        LF_CHECK_ARG(d->ndim == 3 || d->d == 1, "conv: 2-D conv needs d == 1");
        g.kind = 0;
        g.taps = d->ndim == 3 ? d->k * d->k * d->k : d->k * d->k;
        g.out_positions = (int64_t)d->n * d->d * d->h * d->w;
    } else if (d->ndim == 1) {
        g.kind = 1; g.taps = d->d; g.k = d->d;
        g.out_positions = (int64_t)d->n * d->h * d->w;
    } else if (d->ndim == -1) {
        g.kind = 2; g.taps = d->d; g.k = 1;
        g.out_positions = (int64_t)d->n * d->h * d->w;     // per depth slice
    } else {
        LF_CHECK_ARG(false, "conv: ndim %d unsupported", d->ndim);
    }
    return LF_OK;
}

int conv_fp32_launch(const lf_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                     float* rnorm, cudaStream_t st) {
    ConvGeom g;
    if (int e = make_geom(d, g)) return e;
    LF_CHECK_ARG(x && w && y, "conv: null pointer");
    const bool vec = (g.cin % 4) == 0;
    const int64_t mblocks = (g.out_positions + BM - 1) / BM;
    LF_CHECK_ARG(mblocks < (1ll << 31), "conv: too many positions");
    if (g.kind == 2) {
        // expand: act fused; norm (over the whole (d,c) group) runs as a second pass in place
        ConvGeom ge = g; ge.norm = 0;
        if (vec && g.cin <= 32) {
            // depth slices per CTA: amortise the staged input tile, keep >= ~8 CTAs per SM in the grid
            int TS = 4;
            while (TS > 1 && mblocks * ((g.cout + 31) / 32) * ((g.d + TS - 1) / TS) < 8ll * sm_count()) TS /= 2;
            dim3 grid((unsigned)mblocks, (g.cout + 31) / 32, (g.d + TS - 1) / TS);
            expand_multi_kernel<32><<<grid, 128, 0, st>>>(ge, x, w, bias, y, TS, nullptr, nullptr, 0, 0, 1.f);
        } else {
            dim3 grid((unsigned)mblocks, (g.cout + 31) / 32, g.d);
            if (vec) conv_expand_kernel<32, 4><<<grid, 128, 0, st>>>(ge, x, w, bias, y);
            else conv_expand_kernel<32, 1><<<grid, 128, 0, st>>>(ge, x, w, bias, y);
        }
        if (g.norm) {
            const int64_t groups = g.out_positions;
            const int64_t inner = (int64_t)g.h * g.w;
            pixelnorm_group_kernel<<<(unsigned)((groups * 32 + 255) / 256), 256, 0, st>>>(y, y, rnorm, g.n, g.d, inner, g.cout);
        }
        LF_RETURN_LAUNCH();
    }
    if (g.kind == 1 && vec && g.cout <= 32) {
        // depth-collapse: split the depth axis over a cluster of KS CTAs per 128-position tile
        int KS = 8;
        while (KS > 1 && (g.taps % KS) != 0) KS /= 2;       // (a function of the depth only: batch-size independent bits)
        if ((int64_t)g.n * g.d * g.h * g.w * (g.cin / 4) >= (1ll << 32)) KS = 1;
        if (KS > 1) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)mblocks, 1, KS);
            cfg.blockDim = dim3(128);
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = KS;
            cfg.attrs = attr; cfg.numAttrs = 1;
            cudaError_t e = cudaLaunchKernelEx(&cfg, collapse_cluster_kernel, g, x, w, bias, y, rnorm, KS);
            if (e == cudaSuccess) LF_RETURN_LAUNCH();
            // a device / partition that cannot co-schedule the cluster: clear the error and use the one-CTA-per-tile kernel
            (void)cudaGetLastError();
        }
    }
    const bool fuse_norm = g.norm && g.cout <= 64;
    ConvGeom gk = g; gk.norm = fuse_norm ? 1 : 0;
    if (g.cout <= 32) {
        dim3 grid((unsigned)mblocks, 1);
        if (vec) conv_fp32_kernel<32, 4><<<grid, 128, 0, st>>>(gk, x, w, bias, y, rnorm);
        else conv_fp32_kernel<32, 1><<<grid, 128, 0, st>>>(gk, x, w, bias, y, rnorm);
    } else {
        dim3 grid((unsigned)mblocks, (g.cout + 63) / 64);
        if (vec) conv_fp32_kernel<64, 4><<<grid, 128, 0, st>>>(gk, x, w, bias, y, rnorm);
        else conv_fp32_kernel<64, 1><<<grid, 128, 0, st>>>(gk, x, w, bias, y, rnorm);
    }
    if (g.norm && !fuse_norm) {
        pixelnorm_group_kernel<<<(unsigned)((g.out_positions * 32 + 255) / 256), 256, 0, st>>>(
            y, y, rnorm, g.out_positions, 1, 1, g.cout);
    }
    LF_RETURN_LAUNCH();
}

// expand with the fused PixelNorm/LeakyReLU backward epilogue (backward of a depth-collapse whose input came from
// a Block conv): supported for Cin <= 32, Cin % 4 == 0, Cout <= 32, Cout % 4 == 0
int expand_epi_supported(const lf_conv_desc* d) {
    return d->ndim == -1 && d->cin <= 32 && (d->cin & 3) == 0 && d->cout <= 32 && (d->cout & 3) == 0 && !d->norm;
}

int expand_epi_launch(const lf_conv_desc* d, const float* x, const float* w, float* y, const float* epi_y,
                      const float* epi_r, int epi_act, float epi_slope, int epi_norm, cudaStream_t st) {
    ConvGeom g;
    if (int e = make_geom(d, g)) return e;
    LF_CHECK_ARG(expand_epi_supported(d), "expand: fused backward epilogue unsupported for this shape");
    const int64_t mblocks = (g.out_positions + BM - 1) / BM;
    LF_CHECK_ARG(mblocks < (1ll << 31), "conv: too many positions");
    int TS = 4;
    while (TS > 1 && mblocks * ((g.d + TS - 1) / TS) < 8ll * sm_count()) TS /= 2;
    dim3 grid((unsigned)mblocks, 1, (g.d + TS - 1) / TS);
    expand_multi_kernel<32><<<grid, 128, 0, st>>>(g, x, w, nullptr, y, TS, epi_y, epi_r, epi_act, epi_norm, epi_slope);
    LF_RETURN_LAUNCH();
}

}  // namespace lf

using namespace lf;

extern "C" int lf_actnorm_bwd(const float* gy, const float* y, const float* rnorm, float* du,
                              int64_t outer, int gd, int64_t inner, int c, int act, float slope, int norm,
                              void* stream) {
    LF_CHECK_ARG(gy && y && du, "actnorm_bwd: null pointer");
    LF_CHECK_ARG(!norm || rnorm, "actnorm_bwd: norm requires rnorm");
    LF_CHECK_ARG(outer > 0 && gd > 0 && inner > 0 && c > 0, "actnorm_bwd: bad extents");
    const int q4 = c >> 2;
    if (gd == 1 && (c & 3) == 0 && q4 >= 1 && q4 <= 32 && (q4 & (q4 - 1)) == 0) {
        int lg = 0;
        while ((1 << lg) < q4) ++lg;
        const int64_t positions = outer * inner;
        const int64_t units = positions * q4;
        int64_t blocks = (units + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 16;
        if (blocks > cap) blocks = cap;
        actnorm_bwd_vec_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(gy, y, rnorm, du, positions, c, lg,
                                                                                 act, slope, norm);
        LF_RETURN_LAUNCH();
    }
    const int64_t groups = outer * inner;
    actnorm_bwd_kernel<<<(unsigned)((groups * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        gy, y, rnorm, du, outer, gd, inner, c, act, slope, norm);
    LF_RETURN_LAUNCH();
}

extern "C" int lf_conv_bwd_weight(const lf_conv_desc* desc, const float* x, const float* du,
                                  float* grad_w, float* grad_bias, void* stream) {
    ConvGeom g;
    if (int e = make_geom(desc, g)) return e;
    LF_CHECK_ARG(x && du && grad_w, "conv_bwd_weight: null pointer");
    LF_CHECK_ARG((int64_t)g.cin * g.cout <= 16 * 256, "conv_bwd_weight: Cin*Cout = %d > 4096 unsupported",
                 g.cin * g.cout);
    LF_CHECK_ARG(g.cout <= 256, "conv_bwd_weight: Cout > 256 unsupported");
    const size_t smem = (size_t)WG_CHUNK * (g.cin + g.cout) * sizeof(float);
    LF_CHECK_ARG(smem <= 48 * 1024, "conv_bwd_weight: channel counts too large for the staging tile");
    const int64_t target_blocks = (int64_t)sm_count() * 4;
    int64_t ppb = (g.out_positions + target_blocks - 1) / target_blocks;
    ppb = ((ppb + WG_CHUNK - 1) / WG_CHUNK) * WG_CHUNK;
    const int64_t blocks = (g.out_positions + ppb - 1) / ppb;
    conv_bwd_weight_kernel<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(g, x, du, grad_w, grad_bias, ppb);
    LF_RETURN_LAUNCH();
}
