// 3x3x3 Equalized convolution on the 5th-gen tensor cores, "depth-batched" input-stationary formulation (sm_100a).
//
// Reference op: modules/equalized.py:57-64 (conv3d * sqrt(2/fan_in) + bias) + blocks.py:152-158 (LeakyReLU, PixelNorm),
// the camera/object Blocks of recon/models.py — 61 % of a pose-refinement iteration at BASELINE configs[1].
//
// Formulation.  out[d] = sum_dz x[d + dz - 1] * W[dz]: input plane e contributes to the THREE output planes
// e-1, e, e+1 (with W[2], W[1], W[0]).  So instead of re-reading an activation tile from shared memory once per
// filter tap (27 M=128,N=32 MMAs, each moving 4 KB of A for 16 cycles of math: the shared-memory operand port is
// the bottleneck), every (dy,dx) tap pair issues ONE M=128, N=3*Cout MMA whose B operand is [W[2] | W[1] | W[0]]
// and whose accumulator is three neighbouring output planes in TMEM: A is read once per 3 taps.
//   * TMEM holds a ring of S output-plane accumulators per M-tile (NT tiles x S slots x Cout columns <= 512);
//     consecutive planes sit in consecutive column slots, so "planes e-1..e+1" is one contiguous column range
//     (split in two MMAs when the ring wraps); a plane's first contribution is issued with accumulate = 0.
//   * bf16x3 (precision 1): x = x_hi + x_lo, W = W_hi + W_lo in bf16; the three products hi*hi, hi*lo, lo*hi all
//     accumulate into the SAME fp32 TMEM columns (one kernel pass; ~2^-16 relative per product).
//   * Activations arrive in the "split-planar" layout  [part hi|lo][n][d][k-chunk][H+2][W+2][8 x bf16]  with a zero
//     halo: in that layout the input of a run of flattened positions is ONE contiguous range per k-chunk, in exactly
//     the UMMA no-swizzle K-major canonical form (16-byte row per position), so a slab is staged by a handful of
//     1-D bulk copies through the TMA engine (cp.async.bulk, mbarrier tx-count) — no conversion warps, no
//     per-element address math — and a filter tap (dy,dx) is a descriptor start-address offset of (dy*(W+2)+dx)*16 B.
//     M-tiles are 128 consecutive positions of the padded-pitch flattened plane starting at the first interior
//     voxel: ceil(((H-1)(W+2)+W)/128) tiles per plane (33 for 64x64 vs 32 ideal), not rows x ceil(W/128).
//   * Epilogue (8 warps): tcgen05.ld the finished plane, He scale + bias + LeakyReLU + PixelNorm thread-locally
//     (one position = one thread = all Cout channels), then write fp32 channels-last and/or the split-planar form
//     (hi/lo bf16, halo zeros included) for the next convolution.
// Pipelines: slab_full/empty[2] (TMA <-> MMA), acc_full/empty[S] (MMA <-> epilogue); every wait is bounded.
#include "tc_common.cuh"

#include <cuda_bf16.h>

namespace lf {
namespace dz {

using namespace tcx;

constexpr int kThreads = 384;     // warp 0: TMA producer, 1..2: MMA issuers (one per M-tile), 3: TMEM allocator, 4..11: epilogue
constexpr int kRing = 2;
constexpr int kMaxSlots = 8;
constexpr int kSmemBudget = 227 * 1024;

struct Params {
    const uint16_t* x;            // split-planar input, hi part; lo part at x + part_elems
    const uint16_t* wpk;          // packed weights [9][2][KC][3*cout_pad][8]
    const float* bias;
    float* y32;                   // fp32 channels-last output (nullable)
    uint16_t* ysp;                // split-planar output (nullable); lo part at ysp + ypart_elems
    float* rnorm;                 // nullable
    int64_t part_elems, ypart_elems;
    int n, d, h, w, Wp, PP;       // PP = (h+2)*(w+2) positions per padded plane
    int cin_pad, cout, cout_pad, KC, KCo;
    int NT, S, T, NC, DC, ndchunks, items;
    int L, L_alloc;               // slab positions per k-chunk (used / allocated)
    uint32_t slab_bytes, w_bytes;
    int nprod;                    // 3: bf16x3, 1: bf16
    float scale; int act; float slope; int norm;
    // fused backward epilogue (bwd-data feeding a Block conv): the result row g is pushed through the PixelNorm /
    // LeakyReLU backward of the layer that produced this conv's forward input: epi_y = that layer's output in
    // split-planar form (lo part at + epi_part_elems), epi_r = its saved norms; the kernel then writes du_prev
    const uint16_t* epi_y; int64_t epi_part_elems; const float* epi_r; int epi_act, epi_norm; float epi_slope;
    uint64_t magic_Wp;
    long long* dbg;               // diagnostic timeline of CTA 0 (nullable): [role 0..3][event 0..63][2] SM clock stamps
};

__device__ __forceinline__ void stamp(const Params& p, int role, uint32_t idx, int which) {
    if (p.dbg != nullptr && blockIdx.x == 0 && idx < 64) p.dbg[(role * 64 + idx) * 2 + which] = clock64();
}

__device__ __forceinline__ void decode_item(const Params& p, int item, int& n, int& col, int& d0, int& d1) {
    col = item % p.NC; item /= p.NC;
    n = item % p.n;
    const int dchunk = item / p.n;
    d0 = dchunk * p.DC;
    d1 = min(p.d, d0 + p.DC);
}

template <int NCH>      // cout_pad / 16
__global__ void __launch_bounds__(kThreads, 1)
conv3d_dz_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t wsm = smem_u32(smem);
    const uint32_t slabs = wsm + p.w_bytes;
    uint8_t* tail = smem + p.w_bytes + (size_t)kRing * p.slab_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);            // [2] full [2] empty [8] acc_full [8] acc_empty [1] w
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kRing + 2 * kMaxSlots + 1);
    float* bias_s = reinterpret_cast<float*>(bars + 2 * kRing + 2 * kMaxSlots + 2);
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kRing;
    const uint32_t bar_accf = bar_empty + 8 * kRing, bar_acce = bar_accf + 8 * kMaxSlots;
    const uint32_t bar_w = bar_acce + 8 * kMaxSlots;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int n_epi_threads = (p.NT == 2) ? 256 : 128;             // arrivals per acc_empty phase

    if (threadIdx.x == 0) {
        for (int i = 0; i < kRing; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 2); }
        for (int i = 0; i < kMaxSlots; ++i) { mbar_init(bar_accf + 8 * i, 2); mbar_init(bar_acce + 8 * i, n_epi_threads); }
        mbar_init(bar_w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.cout_pad; i += kThreads)
        bias_s[i] = (p.bias != nullptr && i < p.cout) ? p.bias[i] : 0.f;
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t lbo_a = (uint32_t)p.L_alloc * 16u;              // k-chunk stride of a slab
    const uint32_t lbo_b = (uint32_t)(3 * p.cout_pad) * 16u;       // k-chunk stride of a weight block
    const uint32_t part_a = (uint32_t)p.KC * lbo_a;                // hi -> lo inside a slab
    const uint32_t part_b = (uint32_t)p.KC * lbo_b;                // hi -> lo inside a (dy,dx) weight block

    if (warp == 0) {
        // =========================== TMA PRODUCER ===========================
        if (lane == 0) {
            // weights: one contiguous buffer, copied in 16 KB pieces onto one barrier
            mbar_arrive_expect_tx(bar_w, p.w_bytes);
            for (uint32_t off = 0; off < p.w_bytes; off += 16384u) {
                const uint32_t sz = min(16384u, p.w_bytes - off);
                bulk_g2s(wsm + off, reinterpret_cast<const uint8_t*>(p.wpk) + off, sz, bar_w);
            }
            uint32_t ic = 0;
            const int nparts = (p.nprod == 3) ? 2 : 1;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
                int n, col, d0, d1;
                decode_item(p, item, n, col, d0, d1);
                const int e0 = max(d0 - 1, 0), e1 = min(d1, p.d - 1);
                const int start = col * p.NT * 128;
                const int len = min(p.L, p.PP - start);
                const uint32_t bytes_kc = (uint32_t)len * 16u;
                for (int e = e0; e <= e1; ++e, ++ic) {
                    const uint32_t stage = ic % kRing;
                    mbar_wait(bar_empty + 8 * stage, ((ic / kRing) & 1) ^ 1, 1);
                    stamp(p, 0, ic, 0);
                    mbar_arrive_expect_tx(bar_full + 8 * stage, bytes_kc * p.KC * nparts);
                    const uint32_t dst0 = slabs + stage * p.slab_bytes;
                    for (int part = 0; part < nparts; ++part) {
                        const uint16_t* src = p.x + part * p.part_elems +
                                              ((((int64_t)n * p.d + e) * p.KC) * p.PP + start) * 8;
                        for (int kc = 0; kc < p.KC; ++kc)
                            bulk_g2s(dst0 + part * part_a + kc * lbo_a, src + (int64_t)kc * p.PP * 8, bytes_kc,
                                     bar_full + 8 * stage);
                    }
                    stamp(p, 0, ic, 1);
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // =========================== MMA ISSUERS (warp 1: M-tile 0, warp 2: M-tile 1) ===========================
        // One issuer per tile: a single warp needs ~80 cycles of uniform-datapath work per MMA, the tensor pipe 48.
        // The whole warp runs the loop in lock-step so descriptors and TMEM addresses live in UNIFORM registers
        // (UTCHMMA takes uniform operands); only the issue itself is elected.  Everything that varies per plane
        // (column runs of the accumulator ring, B row offsets, instruction descriptors) is computed once per
        // plane; the tap loops only add constants.  Both issuers arrive on every barrier phase (an issuer whose
        // tile lies past the end of the plane issues nothing but keeps pace through the same waits).
        const int my_tile = warp - 1;
        mbar_wait(bar_w, 0, 2);
        tc_fence_after();
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);         // SBO = 128 B (8 contiguous 16-byte rows), version 1
        const uint32_t a_lo_const = (lbo_a >> 4) << 16, b_lo_const = (lbo_b >> 4) << 16;
        const int S = p.S, NT = p.NT, KS = p.cin_pad / 16, NPROD = p.nprod, DEPTH = p.d, T = p.T, NC = p.NC, NB = p.n,
                  DC = p.DC, ITEMS = p.items;
        const uint32_t CP = (uint32_t)p.cout_pad, WP = (uint32_t)p.Wp, SLAB = p.slab_bytes;
        const uint32_t a_ks = 2u * (lbo_a >> 4), b_ks = 2u * (lbo_b >> 4);       // +16 input channels
        const uint32_t a_part = part_a >> 4, b_part = part_b >> 4;               // hi -> lo
        const uint32_t b_tap = 2u * b_part;                                      // next (dy,dx)
        const uint32_t b_base = b_lo_const | (wsm >> 4);
        const uint32_t tbase = tmem_base + (uint32_t)(my_tile * S) * CP;
        uint32_t ic = 0, pcbase = 0;
        for (int item = blockIdx.x; item < ITEMS; item += gridDim.x) {
            int it = item;
            const int col = it % NC; it /= NC;
            const int dchunk = it / NB;
            const int d0 = dchunk * DC, d1 = min(DEPTH, d0 + DC);
            const int e0 = max(d0 - 1, 0), e1 = min(d1, DEPTH - 1);
            const bool active = my_tile < NT && col * NT + my_tile < T;
            uint32_t s_plo = pcbase % S;                            // ring slot of plane max(e-1, d0), kept incrementally
            for (int e = e0; e <= e1; ++e, ++ic) {
                const uint32_t stage = ic % kRing;
                const int plo = max(e - 1, d0), phi = min(e + 1, d1 - 1);
                const int np = phi - plo + 1;
                const int nfresh = ((e + 1 <= phi) ? 1 : 0) + ((e == 0 && plo == 0) ? 1 : 0);
                const uint32_t s0 = s_plo;
                const int nA = min(np, S - (int)s0);               // planes before the ring wraps
                const int nn = np - nfresh;                        // planes already holding partial sums
                const uint32_t rb0 = (uint32_t)(plo - (e - 1));    // B row block of plane plo (0: W[dz=2] .. 2: W[dz=0])
                const uint32_t use0 = (pcbase + (uint32_t)(plo - d0)) / S;
                for (int i = nn; i < np; ++i) {                    // fresh slots must have been drained
                    const uint32_t sl = s0 + (uint32_t)i;
                    const bool wrap = sl >= (uint32_t)S;
                    mbar_wait(bar_acce + 8 * (wrap ? sl - S : sl), ((use0 + (wrap ? 1u : 0u)) & 1) ^ 1, 3);
                }
                mbar_wait(bar_full + 8 * stage, (ic / kRing) & 1, 4);
                tc_fence_after();
                if (lane == 0) stamp(p, 1 + my_tile, ic, 0);
                if (active) {
                    // the (at most two) column runs of this plane set: run 0 = planes [0, nA), run 1 = planes [nA, np)
                    const uint32_t id0 = idesc_bf16((uint32_t)nA * CP), id1 = idesc_bf16((uint32_t)max(np - nA, 1) * CP);
                    const uint32_t brow0 = rb0 * CP, brow1 = (rb0 + (uint32_t)nA) * CP;
                    const bool has1 = np > nA;
                    const uint32_t d_run0 = tbase + s0 * CP, d_run1 = tbase;
                    uint32_t a_row = (a_lo_const | ((slabs + stage * SLAB) >> 4)) + (uint32_t)my_tile * 128u;
                    uint32_t b_cur = b_base;
                    {   // first MMA of the plane set (tap 0, k-step 0, hi*hi): old planes accumulate, fresh ones are overwritten
                        auto emit = [&](int i0, int i1, uint32_t acc) {
                            if (i1 <= i0) return;
                            const uint32_t sl = s0 + (uint32_t)i0;
                            const uint32_t dcol = tbase + (sl >= (uint32_t)S ? sl - S : sl) * CP;
                            const uint32_t b_lo = b_cur + (rb0 + (uint32_t)i0) * CP;
                            if (elect_one()) umma_f16(dcol, a_row, desc_hi, b_lo, desc_hi, idesc_bf16((uint32_t)(i1 - i0) * CP), acc);
                        };
                        emit(0, min(nn, nA), 1u);
                        emit(nA, nn, 1u);
                        emit(nn, nA, 0u);
                        emit(max(nn, nA), np, 0u);
                    }
                    uint32_t skip = 1;                              // the (0,0,0,hi*hi) product was issued above
                    for (int dy = 0; dy < 3; ++dy) {
                        uint32_t a_cur = a_row;
                        for (int dx = 0; dx < 3; ++dx) {
                            uint32_t ak = a_cur, bk = b_cur + brow0;
                            if (!has1) {
                                for (int ks = 0; ks < KS; ++ks) {
                                    const uint32_t ak2 = ak + a_part, bk2 = bk + b_part;
                                    if (elect_one()) {
                                        if (!skip) umma_f16(d_run0, ak, desc_hi, bk, desc_hi, id0, 1u);
                                        if (NPROD == 3) {
                                            umma_f16(d_run0, ak, desc_hi, bk2, desc_hi, id0, 1u);
                                            umma_f16(d_run0, ak2, desc_hi, bk, desc_hi, id0, 1u);
                                        }
                                    }
                                    skip = 0;
                                    ak += a_ks; bk += b_ks;
                                }
                            } else {
                                const uint32_t db = brow1 - brow0;
                                for (int ks = 0; ks < KS; ++ks) {
                                    const uint32_t ak2 = ak + a_part, bk2 = bk + b_part;
                                    if (elect_one()) {
                                        if (!skip) {
                                            umma_f16(d_run0, ak, desc_hi, bk, desc_hi, id0, 1u);
                                            umma_f16(d_run1, ak, desc_hi, bk + db, desc_hi, id1, 1u);
                                        }
                                        if (NPROD == 3) {
                                            umma_f16(d_run0, ak, desc_hi, bk2, desc_hi, id0, 1u);
                                            umma_f16(d_run1, ak, desc_hi, bk2 + db, desc_hi, id1, 1u);
                                            umma_f16(d_run0, ak2, desc_hi, bk, desc_hi, id0, 1u);
                                            umma_f16(d_run1, ak2, desc_hi, bk + db, desc_hi, id1, 1u);
                                        }
                                    }
                                    skip = 0;
                                    ak += a_ks; bk += b_ks;
                                }
                            }
                            a_cur += 1;
                            b_cur += b_tap;
                        }
                        a_row += WP;
                    }
                }
                if (elect_one()) {
                    umma_commit(bar_empty + 8 * stage);                              // slab consumed
                    if (e - 1 >= d0) umma_commit(bar_accf + 8 * s0);                 // plane e-1 (= plo) is complete
                    if (e == DEPTH - 1 && e <= phi) {                                // the last plane of the volume too
                        const uint32_t sl = s0 + (uint32_t)(e - plo);
                        umma_commit(bar_accf + 8 * (sl >= (uint32_t)S ? sl - S : sl));
                    }
                }
                if (lane == 0) stamp(p, 1 + my_tile, ic, 1);
                __syncwarp();
                if (e - 1 >= d0) s_plo = (s_plo + 1 == (uint32_t)S) ? 0u : s_plo + 1;   // plo advances with e from here on
            }
            pcbase += (uint32_t)(d1 - d0);
        }
    } else if (warp >= 4) {
        // =========================== EPILOGUE (2 groups of 4 warps) ===========================
        const int wq = warp & 3;                        // TMEM lane quarter of this warp
        const int grp = (warp - 4) >> 2;
        const float slope = p.act ? p.slope : 1.f;
        const int halo_head = p.Wp + 1;                 // positions [0, halo_head) precede the first interior voxel
        uint32_t pcbase = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            int n, col, d0, d1;
            decode_item(p, item, n, col, d0, d1);
            for (int pl = d0; pl < d1; ++pl) {
                const uint32_t r = pcbase + (uint32_t)(pl - d0);
                const uint32_t slot = r % p.S;
                int tile = grp;
                bool work = true;
                if (p.NT == 1) { tile = 0; work = ((r & 1u) == (uint32_t)grp); }
                if (!work) continue;                    // the other group owns this plane (NT == 1)
                mbar_wait(bar_accf + 8 * slot, (r / p.S) & 1, 5);
                tc_fence_after();
                if (warp == 4 && lane == 0) stamp(p, 3, r, 0);
                const int gt = col * p.NT + tile;       // global tile index inside the plane
                if (gt < p.T) {
                    const int q = halo_head + gt * 128 + wq * 32 + lane;
                    const int yp = fast_div(q, p.magic_Wp), xp = q - yp * p.Wp;
                    const bool valid = (yp >= 1) && (yp <= p.h) && (xp >= 1) && (xp <= p.w);
                    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(tile * p.S + slot) * p.cout_pad;
                    float v[NCH * 16];
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) tmem_ld16(taddr + ch * 16, v + ch * 16);
                    tmem_ld_wait();
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
                        if (valid && p.rnorm != nullptr)
                            p.rnorm[(((int64_t)n * p.d + pl) * p.h + (yp - 1)) * p.w + (xp - 1)] = rn;
                    }
                    if (p.epi_y != nullptr && valid) {
                        // du_prev = gate(y) * (g - y * mean_c(g*y)) / r   with g = this row   (same formula as lf_actnorm_bwd)
                        const uint16_t* yh = p.epi_y + ((((int64_t)n * p.d + pl) * p.KCo) * p.PP + q) * 8;
                        const uint16_t* yl = yh + p.epi_part_elems;
                        float yv[NCH * 16];
#pragma unroll
                        for (int kc = 0; kc < NCH * 2; ++kc) {
                            const uint4 h4 = __ldg(reinterpret_cast<const uint4*>(yh + (int64_t)kc * p.PP * 8));
                            const uint4 l4 = __ldg(reinterpret_cast<const uint4*>(yl + (int64_t)kc * p.PP * 8));
                            const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                yv[kc * 8 + 2 * u] = __uint_as_float(hw[u] << 16) + __uint_as_float(lw[u] << 16);
                                yv[kc * 8 + 2 * u + 1] = __uint_as_float(hw[u] & 0xffff0000u) + __uint_as_float(lw[u] & 0xffff0000u);
                            }
                        }
                        float dot = 0.f, ir = 1.f;
                        if (p.epi_norm) {
#pragma unroll
                            for (int i = 0; i < NCH * 16; ++i) dot += v[i] * yv[i];
                            dot *= 1.f / (float)p.cout;
                            ir = 1.f / __ldg(p.epi_r + (((int64_t)n * p.d + pl) * p.h + (yp - 1)) * p.w + (xp - 1));
                        }
                        const float gs = p.epi_act ? p.epi_slope : 1.f;
#pragma unroll
                        for (int i = 0; i < NCH * 16; ++i) {
                            const float o = (v[i] - yv[i] * dot) * ir;
                            v[i] = yv[i] > 0.f ? o : o * gs;
                        }
                    }
                    if (p.y32 != nullptr && valid) {
                        float* yo = p.y32 + ((((int64_t)n * p.d + pl) * p.h + (yp - 1)) * p.w + (xp - 1)) * p.cout;
#pragma unroll
                        for (int i = 0; i < NCH * 16; i += 4)
                            if (i < p.cout) *reinterpret_cast<float4*>(yo + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    }
                    if (p.ysp != nullptr && q < p.PP) {
                        // split-planar: 16 bytes per (k-chunk, position); halo positions get zeros
                        uint16_t* hi = p.ysp + ((((int64_t)n * p.d + pl) * p.KCo) * p.PP + q) * 8;
                        uint16_t* lo = hi + p.ypart_elems;
#pragma unroll
                        for (int kc = 0; kc < NCH * 2; ++kc) {
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
                tc_fence_before();
                mbar_arrive(bar_acce + 8 * slot);
                if (warp == 4 && lane == 0) stamp(p, 3, r, 1);
                // halo rows of the split-planar output that no tile covers: [0, Wp+1) and [Wp+1 + T*128, PP)
                if (p.ysp != nullptr && tile == 0 && (col == 0 || col == p.NC - 1)) {
                    const int tid = wq * 32 + lane;
                    const int tail0 = halo_head + p.T * 128;
                    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
                    for (int part = 0; part < 2; ++part)
                        for (int kc = 0; kc < p.KCo; ++kc) {
                            uint16_t* base = p.ysp + part * p.ypart_elems + ((((int64_t)n * p.d + pl) * p.KCo + kc) * p.PP) * 8;
                            if (col == 0)
                                for (int q = tid; q < halo_head; q += 128) *reinterpret_cast<uint4*>(base + (int64_t)q * 8) = z4;
                            if (col == p.NC - 1)
                                for (int q = tail0 + tid; q < p.PP; q += 128) *reinterpret_cast<uint4*>(base + (int64_t)q * 8) = z4;
                        }
                }
            }
            pcbase += (uint32_t)(d1 - d0);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 3) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ------------------------------------------------------------------------------------------------------------
// fp32 channels-last [n][d][h][w][c] -> split-planar [hi|lo][n][d][c_pad/8][h+2][w+2][8] bf16 (zero halo, zero padding
// channels).  One thread per padded position: its stores are 16-byte pieces that a warp lays down as 512-byte runs.
__global__ void split_pack_kernel(const float* __restrict__ x, uint16_t* __restrict__ out, int64_t part_elems,
                                  int nd, int h, int w, int c, int KC, uint64_t magic_Wp, uint64_t magic_PP) {
    const int Wp = w + 2, PP = (h + 2) * Wp;
    const int64_t total = (int64_t)nd * PP;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int plane = (int)(idx / PP);
        const int q = (int)(idx - (int64_t)plane * PP);
        const int yp = fast_div(q, magic_Wp), xp = q - yp * Wp;
        const bool valid = yp >= 1 && yp <= h && xp >= 1 && xp <= w;
        const float* src = x + (((int64_t)plane * h + (yp - 1)) * w + (xp - 1)) * c;
        uint16_t* hi = out + ((int64_t)plane * KC * PP + q) * 8;
        uint16_t* lo = hi + part_elems;
        for (int kc = 0; kc < KC; ++kc) {
            uint4 h4 = make_uint4(0u, 0u, 0u, 0u), l4 = h4;
            if (valid) {
                float v[8];
                if ((c & 3) == 0) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (kc * 8 + u * 4 < c) f = ldg4(src + kc * 8 + u * 4);
                        v[u * 4] = f.x; v[u * 4 + 1] = f.y; v[u * 4 + 2] = f.z; v[u * 4 + 3] = f.w;
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = (kc * 8 + u < c) ? __ldg(src + kc * 8 + u) : 0.f;
                }
                split_bf16x2(v[0], v[1], h4.x, l4.x);
                split_bf16x2(v[2], v[3], h4.y, l4.y);
                split_bf16x2(v[4], v[5], h4.z, l4.z);
                split_bf16x2(v[6], v[7], h4.w, l4.w);
            }
            *reinterpret_cast<uint4*>(hi + (int64_t)kc * PP * 8) = h4;
            *reinterpret_cast<uint4*>(lo + (int64_t)kc * PP * 8) = l4;
        }
    }
    (void)magic_PP;
}

// PixelNorm/LeakyReLU backward of a 3-D layer (lf_actnorm_bwd's formula: du = gate(y) * (g - y*mean_c(g*y)) / r) that
// writes its result in split-planar form — what the layer's bwd-data convolution stages with TMA — and, optionally,
// dense fp32 as well (the weight-gradient kernel reads that).  C/4 lanes per position (C in {16, 32}), every lane
// owns 4 channels = 8 bytes of a 16-byte split-planar row; the grid walks PADDED positions so the halo gets its zeros.
__global__ void __launch_bounds__(256)
actnorm_bwd_split_kernel(const float* __restrict__ gy, const float* __restrict__ y, const float* __restrict__ rnorm,
                         float* __restrict__ du, uint16_t* __restrict__ dus, int64_t part_elems, int nd, int h, int w,
                         int c, int lg, int act, float slope, int norm, uint64_t magic_Wp) {
    const int Wp = w + 2, PP = (h + 2) * Wp, q4 = c >> 2, KC = c >> 3;
    const int64_t units = (int64_t)nd * PP * q4;           // multiple of 32? not necessarily: pad for the shuffles
    const int64_t units_pad = (units + 31) & ~(int64_t)31;
    const float inv_c = 1.f / (float)c;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < units_pad; u += (int64_t)gridDim.x * blockDim.x) {
        const bool live = u < units;
        const int64_t uu = live ? u : units - 1;
        const int l = (int)(uu & (q4 - 1));
        const int64_t pidx = uu >> lg;                      // padded position index over all planes
        const int plane = (int)(pidx / PP);
        const int q = (int)(pidx - (int64_t)plane * PP);
        const int yp = fast_div(q, magic_Wp), xp = q - yp * Wp;
        const bool valid = yp >= 1 && yp <= h && xp >= 1 && xp <= w;
        const int64_t pos = valid ? (((int64_t)plane * h + (yp - 1)) * w + (xp - 1)) : 0;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f), yv = g;
        if (valid) { g = ldg4(gy + pos * c + l * 4); yv = ldg4(y + pos * c + l * 4); }
        float4 o = g;
        if (norm) {
            float dot = g.x * yv.x + g.y * yv.y + g.z * yv.z + g.w * yv.w;
            for (int sft = 1; sft < q4; sft <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, sft);
            dot *= inv_c;
            const float ir = valid ? 1.f / __ldg(rnorm + pos) : 0.f;
            o.x = (g.x - yv.x * dot) * ir; o.y = (g.y - yv.y * dot) * ir;
            o.z = (g.z - yv.z * dot) * ir; o.w = (g.w - yv.w * dot) * ir;
        }
        if (act) {
            o.x = yv.x > 0.f ? o.x : o.x * slope; o.y = yv.y > 0.f ? o.y : o.y * slope;
            o.z = yv.z > 0.f ? o.z : o.z * slope; o.w = yv.w > 0.f ? o.w : o.w * slope;
        }
        if (!live) continue;
        if (valid && du != nullptr) *reinterpret_cast<float4*>(du + pos * c + l * 4) = o;
        uint2 hi2, lo2;
        split_bf16x2(o.x, o.y, hi2.x, lo2.x);
        split_bf16x2(o.z, o.w, hi2.y, lo2.y);
        if (!valid) { hi2 = make_uint2(0u, 0u); lo2 = hi2; }
        uint16_t* dst = dus + (((int64_t)plane * KC + (l >> 1)) * PP + q) * 8 + (l & 1) * 4;
        *reinterpret_cast<uint2*>(dst) = hi2;
        *reinterpret_cast<uint2*>(dst + part_elems) = lo2;
    }
}

// fp32 packed weights [27 = (dz,dy,dx)][cin][cout] -> [j = dy*3+dx][part][kc][(2-dz)*cout_pad + co][8 ci] bf16
__global__ void pack_weights_dz_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int cin, int cout,
                                       int cin_pad, int cout_pad) {
    const int KC = cin_pad / 8;
    const int64_t per_j = (int64_t)2 * KC * 3 * cout_pad * 8;
    const int64_t total = 9 * per_j / 2;            // one thread per (j, kc, row, ci) writes hi and lo
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = e;
        const int ci8 = (int)(r % 8); r /= 8;
        const int row = (int)(r % (3 * cout_pad)); r /= 3 * cout_pad;
        const int kc = (int)(r % KC);
        const int j = (int)(r / KC);
        const int rb = row / cout_pad, co = row - rb * cout_pad;
        const int dzi = 2 - rb, ci = kc * 8 + ci8;
        float v = 0.f;
        if (ci < cin && co < cout) v = w[((int64_t)(dzi * 9 + j) * cin + ci) * cout + co];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const int64_t o = (((int64_t)j * 2 * KC + kc) * 3 * cout_pad + row) * 8 + ci8;
        out[o] = __bfloat16_as_ushort(hi);
        out[o + (int64_t)KC * 3 * cout_pad * 8] = __bfloat16_as_ushort(lo);
    }
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

struct Plan {
    int cin_pad, cout_pad, KC, Wp, PP, NT, S, T, NC, DC, ndchunks, L, L_alloc;
    uint32_t slab_bytes, w_bytes, smem_bytes;
};

static bool make_plan(const lf_conv_desc* d, Plan& pl) {
    if (d->ndim != 3 || d->k != 3) return false;
    if (d->precision != 1 && d->precision != 2) return false;
    if (d->n < 1 || d->d < 1 || d->h < 1 || d->w < 1 || d->cin < 1 || d->cout < 1) return false;
    pl.cin_pad = round_up(d->cin, 16);
    pl.cout_pad = round_up(d->cout, 16);
    if (pl.cout_pad > 32 || (d->cout & 3) != 0) return false;       // epilogue keeps the row in registers
    if (3 * pl.cout_pad > 256) return false;
    pl.KC = pl.cin_pad / 8;
    pl.Wp = d->w + 2;
    pl.PP = (d->h + 2) * pl.Wp;
    if (pl.Wp >= 4096 || pl.PP >= (1 << 20)) return false;
    pl.w_bytes = 9u * 2u * pl.KC * 3u * pl.cout_pad * 16u;
    const int span = (d->h - 1) * pl.Wp + d->w;
    pl.T = (span + 127) / 128;
    const uint32_t tail = 8 * (2 * kRing + 2 * kMaxSlots + 2) + 4 * 64 + 128;
    pl.NT = 0;
    for (int nt = 2; nt >= 1; --nt) {
        const int L = nt * 128 + 2 * pl.Wp + 2;
        const int La = round_up(L, 8);
        if ((uint32_t)La * 16u >= (1u << 18)) continue;             // descriptor LBO field: 14 bits of 16-byte units
        const uint32_t slab = 2u * pl.KC * La * 16u;
        if ((uint64_t)pl.w_bytes + (uint64_t)kRing * slab + tail > (uint64_t)kSmemBudget) continue;
        const int S = 512 / (nt * pl.cout_pad);
        if (S < 4) continue;
        pl.NT = nt; pl.L = L; pl.L_alloc = La; pl.slab_bytes = slab; pl.S = S > kMaxSlots ? kMaxSlots : S;
        break;
    }
    if (pl.NT == 0) return false;
    if ((pl.w_bytes >> 4) + (uint32_t)kRing * (pl.slab_bytes >> 4) >= (1u << 14)) return false;   // 14-bit start-address field
    pl.NC = (pl.T + pl.NT - 1) / pl.NT;
    pl.smem_bytes = pl.w_bytes + kRing * pl.slab_bytes + tail;
    // depth chunking: whole columns when they fill the machine, else just enough chunks for one balanced wave
    const int sms = sm_count();
    const int64_t cols = (int64_t)d->n * pl.NC;
    int chunks = 1;
    if (cols * 10 < (int64_t)sms * 9) chunks = (int)((sms + cols - 1) / cols);
    if (chunks > d->d) chunks = d->d;
    pl.DC = (d->d + chunks - 1) / chunks;
    pl.ndchunks = (d->d + pl.DC - 1) / pl.DC;
    return true;
}

}  // namespace dz
}  // namespace lf

using namespace lf;

extern "C" int lf_conv3d_dz_supported(const lf_conv_desc* desc) {
    dz::Plan pl;
    return (desc != nullptr && dz::make_plan(desc, pl)) ? 1 : 0;
}

extern "C" int64_t lf_split_bytes(int n, int d, int h, int w, int c) {
    if (n <= 0 || d <= 0 || h <= 0 || w <= 0 || c <= 0) return 0;
    const int64_t c_pad = (c + 15) / 16 * 16;
    return 2 * (int64_t)n * d * c_pad * (h + 2) * (w + 2) * 2;
}

extern "C" int lf_split_pack(const float* x, void* out, int n, int d, int h, int w, int c, void* stream) {
    LF_CHECK_ARG(x && out && n > 0 && d > 0 && h > 0 && w > 0 && c > 0, "split_pack: bad arguments");
    const int c_pad = (c + 15) / 16 * 16, KC = c_pad / 8, Wp = w + 2;
    const int64_t PP = (int64_t)(h + 2) * Wp;
    LF_CHECK_ARG(PP < (1 << 20) && Wp < 4096, "split_pack: plane too large");
    const int64_t part = (int64_t)n * d * c_pad * PP;
    const int64_t total = (int64_t)n * d * PP;
    const int64_t blocks = (total + 255) / 256;
    dz::split_pack_kernel<<<(unsigned)(blocks > (1 << 20) ? (1 << 20) : blocks), 256, 0, (cudaStream_t)stream>>>(
        x, reinterpret_cast<uint16_t*>(out), part, n * d, h, w, c, KC, tcx::make_magic(Wp), 0);
    LF_RETURN_LAUNCH();
}

extern "C" int64_t lf_conv3d_dz_weight_bytes(int cin, int cout) {
    if (cin <= 0 || cout <= 0) return 0;
    const int64_t cin_pad = (cin + 15) / 16 * 16, cout_pad = (cout + 15) / 16 * 16;
    return 9 * 2 * (cin_pad / 8) * 3 * cout_pad * 16;
}

extern "C" int lf_conv3d_dz_pack_weights(const float* w27, void* out, int cin, int cout, void* stream) {
    LF_CHECK_ARG(w27 && out && cin > 0 && cout > 0, "conv3d_dz_pack_weights: bad arguments");
    const int cin_pad = (cin + 15) / 16 * 16, cout_pad = (cout + 15) / 16 * 16;
    const int64_t total = 9ll * (cin_pad / 8) * 3 * cout_pad * 8;
    dz::pack_weights_dz_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        w27, reinterpret_cast<uint16_t*>(out), cin, cout, cin_pad, cout_pad);
    LF_RETURN_LAUNCH();
}

struct DzEpi { const void* y_split; const float* rnorm; int act, norm; float slope; };

static int conv3d_dz_launch(const lf_conv_desc* desc, const void* x_split, const void* w_packed, const float* bias,
                            float* y32, void* y_split, float* rnorm, long long* dbg, void* stream,
                            const DzEpi* epi = nullptr) {
    dz::Plan pl;
    if (desc == nullptr || !dz::make_plan(desc, pl)) {
        set_error("conv3d_dz: unsupported shape/precision (needs 3-D k=3, Cout in {4..32, %%4}, precision 1|2)");
        return LF_EUNSUPPORTED;
    }
    LF_CHECK_ARG(x_split && w_packed && (y32 || y_split), "conv3d_dz: null pointer");
    dz::Params p;
    p.x = reinterpret_cast<const uint16_t*>(x_split);
    p.wpk = reinterpret_cast<const uint16_t*>(w_packed);
    p.bias = bias; p.y32 = y32; p.ysp = reinterpret_cast<uint16_t*>(y_split); p.rnorm = rnorm;
    p.n = desc->n; p.d = desc->d; p.h = desc->h; p.w = desc->w; p.Wp = pl.Wp; p.PP = pl.PP;
    p.cin_pad = pl.cin_pad; p.cout = desc->cout; p.cout_pad = pl.cout_pad; p.KC = pl.KC; p.KCo = pl.cout_pad / 8;
    p.part_elems = (int64_t)desc->n * desc->d * pl.cin_pad * pl.PP;
    p.ypart_elems = (int64_t)desc->n * desc->d * pl.cout_pad * pl.PP;
    p.NT = pl.NT; p.S = pl.S; p.T = pl.T; p.NC = pl.NC; p.DC = pl.DC; p.ndchunks = pl.ndchunks;
    p.items = desc->n * pl.NC * pl.ndchunks;
    p.L = pl.L; p.L_alloc = pl.L_alloc; p.slab_bytes = pl.slab_bytes; p.w_bytes = pl.w_bytes;
    p.nprod = desc->precision == 1 ? 3 : 1;
    p.scale = desc->scale; p.act = desc->act; p.slope = desc->slope; p.norm = desc->norm;
    p.magic_Wp = tcx::make_magic(pl.Wp);
    p.dbg = dbg;
    p.epi_y = epi ? reinterpret_cast<const uint16_t*>(epi->y_split) : nullptr;
    p.epi_part_elems = p.ypart_elems;
    p.epi_r = epi ? epi->rnorm : nullptr; p.epi_act = epi ? epi->act : 0; p.epi_norm = epi ? epi->norm : 0;
    p.epi_slope = epi ? epi->slope : 1.f;
    void (*kern)(dz::Params) = pl.cout_pad == 16 ? dz::conv3d_dz_kernel<1> : dz::conv3d_dz_kernel<2>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dz::kSmemBudget);
    if (e != cudaSuccess) { set_error("conv3d_dz: cannot raise dynamic smem: %s", cudaGetErrorString(e)); return (int)e; }
    const int grid = p.items < sm_count() ? p.items : sm_count();
    kern<<<grid, dz::kThreads, pl.smem_bytes, (cudaStream_t)stream>>>(p);
    LF_RETURN_LAUNCH();
}

extern "C" int lf_conv3d_dz(const lf_conv_desc* desc, const void* x_split, const void* w_packed, const float* bias,
                            float* y32, void* y_split, float* rnorm, void* stream) {
    return conv3d_dz_launch(desc, x_split, w_packed, bias, y32, y_split, rnorm, nullptr, stream);
}

// Diagnostic variant: additionally records SM-clock stamps of CTA 0's pipeline roles into `stamps`
// (device buffer of 4*64*2 int64: [producer | issuer 0 | issuer 1 | epilogue group 0][event][begin, end]).
extern "C" int lf_conv3d_dz_timeline(const lf_conv_desc* desc, const void* x_split, const void* w_packed,
                                     const float* bias, float* y32, void* y_split, float* rnorm, void* stamps,
                                     void* stream) {
    LF_CHECK_ARG(stamps, "conv3d_dz_timeline: null stamp buffer");
    return conv3d_dz_launch(desc, x_split, w_packed, bias, y32, y_split, rnorm, reinterpret_cast<long long*>(stamps), stream);
}

// bwd-data convolution (desc: cin = forward Cout, cout = forward Cin, act = norm = 0, no bias) whose result row is pushed
// through the PixelNorm/LeakyReLU backward of the layer that produced the forward input: writes du_prev (dense fp32
// and/or split-planar).  y_prev_split is that layer's output in split-planar form, rnorm_prev its saved norms.
extern "C" int lf_conv3d_dz_bwd_epi(const lf_conv_desc* desc, const void* du_split, const void* w_packed,
                                    const void* y_prev_split, const float* rnorm_prev, int prev_act, float prev_slope,
                                    int prev_norm, float* du_prev32, void* du_prev_split, void* stream) {
    LF_CHECK_ARG(desc && !desc->act && !desc->norm, "conv3d_dz_bwd_epi: the bwd-data descriptor has no activation/norm of its own");
    LF_CHECK_ARG(y_prev_split && (!prev_norm || rnorm_prev), "conv3d_dz_bwd_epi: null pointer");
    DzEpi epi{y_prev_split, rnorm_prev, prev_act, prev_norm, prev_slope};
    return conv3d_dz_launch(desc, du_split, w_packed, nullptr, du_prev32, du_prev_split, nullptr, nullptr, stream, &epi);
}

extern "C" int lf_actnorm_bwd_split(const float* gy, const float* y, const float* rnorm, float* du, void* du_split,
                                    int n, int d, int h, int w, int c, int act, float slope, int norm, void* stream) {
    LF_CHECK_ARG(gy && y && du_split && (!norm || rnorm), "actnorm_bwd_split: null pointer");
    LF_CHECK_ARG(n > 0 && d > 0 && h > 0 && w > 0 && (c == 16 || c == 32), "actnorm_bwd_split: C must be 16 or 32");
    const int Wp = w + 2;
    const int64_t PP = (int64_t)(h + 2) * Wp;
    LF_CHECK_ARG(PP < (1 << 20) && Wp < 4096, "actnorm_bwd_split: plane too large");
    const int lg = (c == 32) ? 3 : 2;
    const int64_t units = (int64_t)n * d * PP * (c >> 2);
    int64_t blocks = (units + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    dz::actnorm_bwd_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        gy, y, rnorm, du, reinterpret_cast<uint16_t*>(du_split), (int64_t)n * d * c * PP, n * d, h, w, c, lg, act, slope,
        norm, tcx::make_magic(Wp));
    LF_RETURN_LAUNCH();
}
