// Weight gradient of the 3x3x3 Equalized convolution on the 5th-gen tensor cores (sm_100a).
//
// Reference op: the autograd of modules/equalized.py:57-64 inside ReconTrainer.run_iteration
// (tools/train/train_reconstruct.py:523-534) — 90 % of a training iteration while it ran on the FFMA kernel.
//
//   dW[dz][dy][dx][ci][co] = sum over (n, d, h, w) of  x[n, d+dz-1, h+dy-1, w+dx-1, ci] * du[n, d, h, w, co]
//
// GEMM view: M = ci, N = co, K = positions — tiny M and N, enormous K.  Both operands arrive in the library's
// split-planar layout [hi|lo][n][d][c/8][(H+2)(W+2)][8 x bf16] (zero halo), i.e. with K (the flattened padded
// position) strided by 16 bytes and 8 channels contiguous: that is exactly the UMMA *MN-major* no-swizzle canonical
// form (core matrix = 8 positions x 8 channels = 128 contiguous bytes, LBO = 128 B between k-groups, SBO = the
// stride between 8-channel groups), so a run of positions is staged by plain 1-D bulk TMA copies and consumed
// without any transposition.  A filter tap is a start-address offset of the x operand relative to the du operand;
// halo positions hold zeros in both, so the sum simply runs over the interior span of the flattened padded plane.
//
// To give the instruction a useful shape the M dimension is filled with
//     [x_hi(ci) ; x_lo(ci)]  x  NCOPY copies of the same run shifted by +1 position (consecutive dx taps)
// and N with [du_hi(co) | du_lo(co)]: ONE M=128 MMA yields all four hi/lo products of up to four dx taps (the row
// and column halves are added in the reduction; hi*hi + hi*lo + lo*hi + lo*lo is the exact product of the bf16x2
// splits, fp32 accumulation in TMEM).  A CTA owns one dy tap (dyt = blockIdx % 3: the x run is offset by
// (dyt-1) rows), all three dz (x plane e pairs with du planes e+1, e, e-1: a ring of du planes, 3 x stages) and a
// strided set of (sample, 128-position chunk, depth segment) columns that it marches through in depth; its 3*NM
// accumulators (<= 384 TMEM columns) live across all its items and are written once, as per-CTA partials that a second
// kernel sums in a fixed order (deterministic, no atomics).  One MMA-issuer warp per dz (a warp sustains one tcgen05.mma
// per ~120 cycles), one bulk copy per producer lane.  The same kernel takes 2-D 3x3 layers (one plane per image, the
// centre dz only, up to 64 channels: the accumulator budget is then NM * N <= 512 columns).
#include "tc_common.cuh"

#include <cuda_bf16.h>

namespace lf {
namespace dw {

using namespace tcx;

constexpr int kThreads = 256;      // warp 0: TMA producer, 1..3: MMA issuers (one per dz; warp 2 also allocates TMEM), 4..7: epilogue
constexpr int kXStages = 3;
constexpr int kDyRing = 6;        // du planes e-1, e, e+1 of the step in flight + three prefetched (Params::dring <= this)
constexpr int kChunk = 128;        // positions per item step (8 k-steps of 16)

struct Params {
    const uint16_t* x;             // split-planar forward input (hi part; lo part at + x_part)
    const uint16_t* dy;            // split-planar gradient of the pre-activation output
    int64_t x_part, dy_part;       // elements per part
    const uint16_t* x_end;         // one past the last element that may be read (copies are clamped to it)
    const uint16_t* dy_end;
    float* ws;                     // [ctas][NACC][128][N] partial accumulators
    int n, d, Wp, PP, KCi, KCo, nparts;
    int NCOPY, NM, R, N;           // x copies stacked in M, MMAs per dz, rows per copy, columns
    int span, nchunks, items, G;   // interior positions per plane, 128-position chunks, (sample, chunk, depth segment) items, CTAs per dy tap
    int nseg, dseg;                // depth segments per column and planes per segment
    int two_d, dring;              // 2-D convolution (one plane per image, dz = 1 only); du ring slots in use
    int x_kcs, x_kc0, dy_kcs, dy_kc0;   // 8-channel chunks per plane of the whole buffers, first chunk of this call's channel group
    int Lx;                        // positions per x region (allocated)
    uint32_t x_stage_bytes, dy_slot_bytes;
};

// instruction descriptor: kind::f16, D = F32, A = B = BF16, BOTH MN-major (bits 15, 16), M = 128
__device__ __forceinline__ uint32_t idesc_mn(uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__global__ void __launch_bounds__(kThreads, 1)
conv3d_dw_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t xs0 = smem_u32(smem);
    const uint32_t dys0 = xs0 + kXStages * p.x_stage_bytes;
    uint8_t* tail = smem + (size_t)kXStages * p.x_stage_bytes + (size_t)p.dring * p.dy_slot_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);     // full_x[2] empty_x[2] full_dy[4] empty_dy[4] done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kXStages + 2 * kDyRing + 1);
    uint32_t* touched_s = tmem_slot + 1;                     // [3]: accumulators each issuer has written
    const uint32_t bar_fx = smem_u32(bars), bar_ex = bar_fx + 8 * kXStages;
    const uint32_t bar_fd = bar_ex + 8 * kXStages, bar_ed = bar_fd + 8 * kDyRing, bar_done = bar_ed + 8 * kDyRing;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int dyt = blockIdx.x % 3, slot = blockIdx.x / 3;

    // stale bytes behind a copy clamped at the end of a buffer must be finite: they meet halo zeros of the other operand
    for (uint32_t i = threadIdx.x * 16; i < kXStages * p.x_stage_bytes + p.dring * p.dy_slot_bytes; i += kThreads * 16)
        *reinterpret_cast<uint4*>(smem + i) = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) {
        for (int i = 0; i < kXStages; ++i) { mbar_init(bar_fx + 8 * i, 1); mbar_init(bar_ex + 8 * i, 3); }
        for (int i = 0; i < kDyRing; ++i) { mbar_init(bar_fd + 8 * i, 1); mbar_init(bar_ed + 8 * i, 3); }
        mbar_init(bar_done, 3);
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

    const int NACC = p.two_d ? p.NM : 3 * p.NM;
    const uint32_t DR = (uint32_t)p.dring;
    const uint32_t x_region = (uint32_t)p.Lx * 16u;                 // bytes per 8-channel group of one copy
    const int xg = p.R / 8;                                         // 8-channel groups per copy (parts x chunks)
    const int dg = p.N / 8;
    const int first = p.Wp + 1;                                     // first interior position of a padded plane

    if (warp == 0) {
        // =========================== TMA PRODUCER ===========================
        // One bulk copy per lane: a single thread needs ~150 cycles to issue each cp.async.bulk, and a step has 24 of
        // them (16 x regions + 8 du regions of ~2 KB) against ~2.3k cycles of MMA work.  Lanes 0..15 own the x regions of
        // the step's stage, lanes 16..16+dg-1 the du regions of the next plane; lane 0 posts the byte counts.
        uint32_t sx = 0, pc = 0;                                 // x stage counter, du plane counter
        for (int item = slot; item < p.items; item += p.G) {
            const int seg = item % p.nseg, col = item / p.nseg;
            const int n = col / p.nchunks, c = col - n * p.nchunks;
            const int e0 = seg * p.dseg, e1 = min(p.d, e0 + p.dseg);      // x planes of this item
            const int dlo = max(e0 - 1, 0), dhi = min(e1, p.d - 1);       // du planes it touches
            const int q0 = first + c * kChunk;
            const int ksteps = min(kChunk / 16, (p.span - c * kChunk + 15) / 16);
            const uint32_t dy_bytes = (uint32_t)ksteps * 256u;
            const uint32_t x_bytes = (uint32_t)(ksteps * 16 + (p.NM - 1) * p.NCOPY) * 16u;
            const bool x_lane = lane < p.NCOPY * xg, d_lane = lane >= 16 && lane < 16 + dg;
            // this lane's x region: copy cp, channel group g (part, kc)
            int64_t x_off = 0;
            if (x_lane) {
                const int cp = lane / xg, g = lane - cp * xg;
                const int part = g / p.KCi, kc = g - part * p.KCi;
                x_off = part * p.x_part + ((((int64_t)n * p.d) * p.x_kcs + p.x_kc0 + kc) * p.PP + q0 + (dyt - 1) * p.Wp - 1 + cp) * 8;
            }
            int64_t d_off = 0;
            if (d_lane) {
                const int g = lane - 16;
                const int part = g / p.KCo, kc = g - part * p.KCo;
                d_off = part * p.dy_part + ((((int64_t)n * p.d) * p.dy_kcs + p.dy_kc0 + kc) * p.PP + q0) * 8;
            }
            const int64_t x_plane = (int64_t)p.x_kcs * p.PP * 8, d_plane = (int64_t)p.dy_kcs * p.PP * 8;
            auto load_dy = [&](int dpl) {
                const uint32_t r = pc + (uint32_t)(dpl - dlo), sl = r % DR;
                const uint16_t* dsrc = p.dy + d_off + dpl * d_plane;
                uint32_t dbytes = 0;
                if (d_lane) {            // a narrow plane's rounded-up run may pass the end of the buffer: clamp
                    const int64_t avail = (p.dy_end - dsrc) * 2;
                    dbytes = avail <= 0 ? 0u : (uint32_t)min((int64_t)dy_bytes, avail);
                }
                const uint32_t dtotal = __reduce_add_sync(0xffffffffu, dbytes);
                if (lane == 0) {
                    mbar_wait(bar_ed + 8 * sl, ((r / DR) & 1) ^ 1, 11);
                    mbar_arrive_expect_tx(bar_fd + 8 * sl, dtotal);
                }
                __syncwarp();
                if (dbytes)
                    bulk_g2s(dys0 + sl * p.dy_slot_bytes + (uint32_t)(lane - 16) * (kChunk * 16), dsrc, dbytes, bar_fd + 8 * sl);
            };
            for (int dpl = dlo; dpl <= e0; ++dpl) load_dy(dpl);
            for (int e = e0; e < e1; ++e, ++sx) {
                const uint32_t st = sx % kXStages;
                const uint16_t* src = p.x + x_off + e * x_plane;
                uint32_t bytes = 0;
                if (x_lane) {
                    const int64_t avail = (p.x_end - src) * 2;       // the very last region may end at the buffer end
                    bytes = avail <= 0 ? 0u : (uint32_t)min((int64_t)x_bytes, avail);
                }
                const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
                if (lane == 0) {
                    mbar_wait(bar_ex + 8 * st, ((sx / kXStages) & 1) ^ 1, 12);
                    mbar_arrive_expect_tx(bar_fx + 8 * st, total);
                }
                __syncwarp();
                if (bytes) bulk_g2s(xs0 + st * p.x_stage_bytes + (uint32_t)lane * x_region, src, bytes, bar_fx + 8 * st);
                if (e + 1 <= dhi) load_dy(e + 1);
            }
            pc += (uint32_t)(dhi - dlo + 1);
        }
    } else if (warp <= 3) {
        // =========================== MMA ISSUERS (warp 1 + dz) ===========================
        // A warp issues one tcgen05.mma per ~120 cycles (uniform-datapath work + issue latency), the tensor pipe takes
        // one of these every ~48: three issuers, one per dz (its own accumulators, its own du plane of the ring).
        const int iz = warp - 1;
        const int dz = p.two_d ? 1 : iz;                             // 2-D: one plane per image, the centre dz only
        const int j0 = p.two_d ? iz : 0, j1 = p.two_d ? min(iz + 1, p.NM) : p.NM;   // 2-D: issuer iz owns accumulator iz
        const uint32_t idesc = idesc_mn((uint32_t)p.N);
        const uint32_t a_hi = (uint32_t)p.Lx | (1u << 14);           // SBO = one x region (16-byte units), version 1
        const uint32_t b_hi = (uint32_t)kChunk | (1u << 14);         // SBO = one du region
        const uint32_t lbo = 8u << 16;                               // LBO = 128 bytes: the next 8 positions
        uint32_t touched = 0, sx = 0, pc = 0;
        for (int item = slot; item < p.items; item += p.G) {
            const int seg = item % p.nseg, c = (item / p.nseg) % p.nchunks;
            const int e0 = seg * p.dseg, e1 = min(p.d, e0 + p.dseg);
            const int dlo = max(e0 - 1, 0), dhi = min(e1, p.d - 1);
            const int ksteps = min(kChunk / 16, (p.span - c * kChunk + 15) / 16);
            for (int e = e0; e < e1; ++e, ++sx) {
                const uint32_t st = sx % kXStages;
                const int dpl = e - dz + 1;                          // du plane paired with x plane e for this dz
                const bool live = dpl >= 0 && dpl < p.d && j0 < j1;
                if (live) {
                    const uint32_t r = pc + (uint32_t)(dpl - dlo);
                    mbar_wait(bar_fd + 8 * (r % DR), (r / DR) & 1, 13);
                }
                mbar_wait(bar_fx + 8 * st, (sx / kXStages) & 1, 15);
                tc_fence_after();
                if (live) {
                    const uint32_t a0 = lbo | ((xs0 + st * p.x_stage_bytes) >> 4);
                    const uint32_t b0 = lbo | ((dys0 + ((pc + (uint32_t)(dpl - dlo)) % DR) * p.dy_slot_bytes) >> 4);
                    for (int ks = 0; ks < ksteps; ++ks) {
                        for (int j = j0; j < j1; ++j) {
                            const int acc = p.two_d ? j : dz * p.NM + j;
                            if (elect_one())
                                umma_f16(tmem_base + (uint32_t)(acc * p.N), a0 + (uint32_t)(ks * 16 + j * p.NCOPY), a_hi,
                                         b0 + (uint32_t)ks * 16u, b_hi, idesc, (touched >> acc) & 1u);
                            touched |= 1u << acc;
                        }
                    }
                }
                if (elect_one()) {
                    umma_commit(bar_ex + 8 * st);
                    // a du plane is last used by the step of x plane (plane + 1), or by the item's last step
                    if (e - 1 >= dlo) umma_commit(bar_ed + 8 * ((pc + (uint32_t)(e - 1 - dlo)) % DR));
                    if (e == e1 - 1)
                        for (int q = e; q <= dhi; ++q) umma_commit(bar_ed + 8 * ((pc + (uint32_t)(q - dlo)) % DR));
                }
                __syncwarp();
            }
            pc += (uint32_t)(dhi - dlo + 1);
        }
        if (lane == 0) { touched_s[iz] = touched; __threadfence_block(); }
        __syncwarp();
        if (elect_one()) umma_commit(bar_done);
    } else if (warp >= 4) {
        // =========================== EPILOGUE: TMEM -> per-CTA partials ===========================
        const int wq = warp & 3;
        mbar_wait(bar_done, 0, 16);
        tc_fence_after();
        const volatile uint32_t* tv = touched_s;
        const uint32_t touched = tv[0] | tv[1] | tv[2];
        const int row = wq * 32 + lane;
        float* out = p.ws + ((int64_t)blockIdx.x * NACC * 128 + row) * p.N;
        for (int acc = 0; acc < NACC; ++acc) {
            float* o = out + (int64_t)acc * 128 * p.N;
            const bool live = (touched >> acc) & 1u;
            for (int c0 = 0; c0 < p.N; c0 += 16) {
                float v[16];
                if (live) {
                    tmem_ld16(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * p.N + c0), v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    *reinterpret_cast<float4*>(o + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// partials [G*3 ctas][3*NM][128][N] -> grad_w [27][cin][cout]: fixed summation order (CTA slot, then x part, then du part)
__global__ void dw_reduce_kernel(const float* __restrict__ ws, float* __restrict__ gw, int G, int NM, int NCOPY, int R, int N,
                                 int cin, int cout, int cin_pad, int cout_pad, int nparts, float scale, int two_d,
                                 int cin_total, int cout_total, int ci_off, int co_off) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (two_d ? 9 : 27) * cin * cout) return;
    const int co = idx % cout, ci = (idx / cout) % cin, tap = idx / (cout * cin);
    const int dx = tap % 3, dyt = (tap / 3) % 3, dz = two_d ? 0 : tap / 9;
    const int j = dx / NCOPY, cp = dx - j * NCOPY;
    const int NACC = two_d ? NM : 3 * NM;
    float s = 0.f;
    for (int g = 0; g < G; ++g) {
        const float* base = ws + (((int64_t)(g * 3 + dyt) * NACC + dz * NM + j) * 128) * N;
        for (int px = 0; px < nparts; ++px)
            for (int py = 0; py < nparts; ++py)
                s += base[(int64_t)(cp * R + px * cin_pad + ci) * N + py * cout_pad + co];
    }
    // du is the gradient of conv(x, W * he): d/dW carries the He constant
    gw[((int64_t)tap * cin_total + ci_off + ci) * cout_total + co_off + co] = s * scale;
}

// bias gradient: per-channel sum of a split-planar volume (hi + lo), two fixed-order stages
__global__ void __launch_bounds__(256)
split_colsum_kernel(const uint16_t* __restrict__ v, int64_t part_elems, int nparts, int KC, int PP, float* __restrict__ partial) {
    const int plane = blockIdx.x;                      // (n, d)
    __shared__ float red[8][8];
    for (int kc = 0; kc < KC; ++kc) {
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int part = 0; part < nparts; ++part) {
            const uint16_t* src = v + part * part_elems + ((int64_t)plane * KC + kc) * PP * 8;
            for (int q = threadIdx.x; q < PP; q += 256) {
                const uint4 w4 = __ldg(reinterpret_cast<const uint4*>(src + (int64_t)q * 8));
                const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    a[2 * u] += __uint_as_float(w[u] << 16);
                    a[2 * u + 1] += __uint_as_float(w[u] & 0xffff0000u);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float s = warp_sum(a[i]);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = s;
        }
        __syncthreads();
        if (threadIdx.x < 8) {
            float s = 0.f;
            for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
            partial[(int64_t)plane * KC * 8 + kc * 8 + threadIdx.x] = s;
        }
        __syncthreads();
    }
}

__global__ void colsum_finish_kernel(const float* __restrict__ partial, int planes, int cpad, int c, float* __restrict__ out) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    float s = 0.f;
    for (int pl = 0; pl < planes; ++pl) s += partial[(int64_t)pl * cpad + ch];
    out[ch] = s;
}

struct Plan {
    int cin_pad, cout_pad, KCi, KCo, nparts, NCOPY, NM, R, N, Wp, PP, span, nchunks, items, G, Lx, nseg, dseg, two_d, dring, nacc;
    uint32_t x_stage_bytes, dy_slot_bytes, smem_bytes;
};

static bool make_plan(const lf_conv_desc* d, Plan& pl) {
    if (d->k != 3 || !(d->ndim == 3 || (d->ndim == 2 && d->d == 1))) return false;
    pl.two_d = d->ndim == 2;
    if (d->precision != 1 && d->precision != 2) return false;
    if (d->n < 1 || d->d < 1 || d->h < 1 || d->w < 1 || d->cin < 1 || d->cout < 1) return false;
    pl.cin_pad = (d->cin + 15) / 16 * 16;
    pl.cout_pad = (d->cout + 15) / 16 * 16;
    pl.nparts = d->precision == 1 ? 2 : 1;
    pl.R = pl.nparts * pl.cin_pad;
    pl.N = pl.nparts * pl.cout_pad;
    if (pl.R != 32 && pl.R != 64 && !(pl.two_d && pl.R == 128)) return false;   // M = NCOPY * R = 128, NCOPY in {4, 2, 1}
    pl.NCOPY = 128 / pl.R;
    pl.NM = (3 + pl.NCOPY - 1) / pl.NCOPY;
    pl.nacc = pl.two_d ? pl.NM : 3 * pl.NM;
    if (pl.N > 128 || pl.nacc * pl.N > 512) return false;       // accumulators live in TMEM (512 columns)
    pl.KCi = pl.cin_pad / 8; pl.KCo = pl.cout_pad / 8;
    pl.Wp = d->w + 2;
    pl.PP = (d->h + 2) * pl.Wp;
    if (pl.Wp >= 4096 || pl.PP >= (1 << 20)) return false;
    pl.span = (d->h - 1) * pl.Wp + d->w;
    pl.nchunks = (pl.span + kChunk - 1) / kChunk;
    const int per_tap = sm_count() / 3;
    // few columns (the GRU's 2-object batches): cut the depth march into segments (one extra du plane at each cut)
    pl.nseg = 1;
    while (pl.nseg < 8 && (int64_t)d->n * pl.nchunks * pl.nseg < 3ll * per_tap && d->d / (pl.nseg * 2) >= 8) pl.nseg *= 2;
    pl.dseg = (d->d + pl.nseg - 1) / pl.nseg;
    pl.nseg = (d->d + pl.dseg - 1) / pl.dseg;
    const int64_t items = (int64_t)d->n * pl.nchunks * pl.nseg;
    if (items >= (1ll << 30)) return false;
    pl.items = (int)items;
    pl.G = pl.items < per_tap ? pl.items : per_tap;
    pl.Lx = kChunk + (pl.NM - 1) * pl.NCOPY;
    pl.Lx = (pl.Lx + 7) / 8 * 8;
    pl.x_stage_bytes = (uint32_t)(pl.NCOPY * (pl.R / 8)) * pl.Lx * 16u;
    pl.dy_slot_bytes = (uint32_t)(pl.N / 8) * kChunk * 16u;
    const uint32_t tail = 8 * (2 * kXStages + 2 * kDyRing + 1) + 64;
    for (pl.dring = kDyRing; pl.dring >= (pl.two_d ? 2 : 4); --pl.dring) {
        pl.smem_bytes = kXStages * pl.x_stage_bytes + pl.dring * pl.dy_slot_bytes + tail;
        if (pl.smem_bytes <= 227u * 1024u) return true;
    }
    return false;
}

}  // namespace dw
}  // namespace lf

using namespace lf;

// Layers wider than one call's operand shapes are evaluated as channel-group pairs (32 input x 32 output channels per
// call, each reading its chunks of the same split-planar buffers); `sub` is the descriptor of one pair.
static bool plan_for(const lf_conv_desc* desc, dw::Plan& pl, lf_conv_desc& sub, int& gin, int& gout) {
    if (desc == nullptr) return false;
    sub = *desc; gin = gout = 1;
    if (dw::make_plan(desc, pl)) return true;
    if (desc->precision != 1 && desc->precision != 2) return false;
    const int cin_pad = (desc->cin + 15) / 16 * 16, cout_pad = (desc->cout + 15) / 16 * 16;
    // (a 16-channel tail group reads two chunks past its plane: finite data of the next plane, or the zero-filled stage
    //  behind a copy clamped at the end of the buffer; those rows / columns of the accumulators are never read)
    sub.cin = 32; sub.cout = 32;
    if (!dw::make_plan(&sub, pl)) return false;
    gin = (cin_pad + 31) / 32; gout = (cout_pad + 31) / 32;
    return true;
}

extern "C" int lf_conv3d_dw_supported(const lf_conv_desc* desc) {
    dw::Plan pl; lf_conv_desc sub; int gi, go;
    return plan_for(desc, pl, sub, gi, go) ? 1 : 0;
}

// workspace (floats): per-CTA partial accumulators + the bias column-sum partials
extern "C" int64_t lf_conv3d_dw_ws(const lf_conv_desc* desc) {
    dw::Plan pl; lf_conv_desc sub; int gi, go;
    if (!plan_for(desc, pl, sub, gi, go)) return 0;
    const int cout_pad = (desc->cout + 15) / 16 * 16;
    return (int64_t)pl.G * 3 * pl.nacc * 128 * pl.N + (int64_t)desc->n * desc->d * cout_pad;
}

// grad_w_packed [27 | 9][Cin][Cout] (overwritten) = desc->scale * sum x (.) du over all positions; grad_bias [Cout]
// (nullable) = sum du.  x_split / du_split: split-planar volumes of the forward input and of
// d(loss)/d(pre-activation output) (precision 1: hi and lo parts; 2: hi parts only).
// Reference: autograd of modules/equalized.py:57-64.
extern "C" int lf_conv3d_dw(const lf_conv_desc* desc, const void* x_split, const void* du_split, float* ws,
                            float* grad_w_packed, float* grad_bias, void* stream) {
    dw::Plan pl; lf_conv_desc sub; int gin, gout;
    if (!plan_for(desc, pl, sub, gin, gout)) {
        set_error("conv3d_dw: unsupported shape/precision (k=3; one call: parts*Cin_pad in {32, 64[, 128 in 2-D]}; wider "
                  "layers: 32x32 channel-group pairs)");
        return LF_EUNSUPPORTED;
    }
    LF_CHECK_ARG(x_split && du_split && ws && grad_w_packed, "conv3d_dw: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int cin_pad = (desc->cin + 15) / 16 * 16, cout_pad = (desc->cout + 15) / 16 * 16;
    dw::Params p;
    p.x = reinterpret_cast<const uint16_t*>(x_split);
    p.dy = reinterpret_cast<const uint16_t*>(du_split);
    p.x_part = (int64_t)desc->n * desc->d * cin_pad * pl.PP;
    p.dy_part = (int64_t)desc->n * desc->d * cout_pad * pl.PP;
    p.x_end = p.x + 2 * p.x_part;            // the buffers always hold both parts (lf_split_bytes)
    p.dy_end = p.dy + 2 * p.dy_part;
    p.ws = ws;
    p.n = desc->n; p.d = desc->d; p.Wp = pl.Wp; p.PP = pl.PP; p.KCi = pl.KCi; p.KCo = pl.KCo; p.nparts = pl.nparts;
    p.NCOPY = pl.NCOPY; p.NM = pl.NM; p.R = pl.R; p.N = pl.N;
    p.span = pl.span; p.nchunks = pl.nchunks; p.items = pl.items; p.G = pl.G; p.Lx = pl.Lx;
    p.nseg = pl.nseg; p.dseg = pl.dseg; p.two_d = pl.two_d; p.dring = pl.dring;
    p.x_kcs = cin_pad / 8; p.dy_kcs = cout_pad / 8;
    p.x_stage_bytes = pl.x_stage_bytes; p.dy_slot_bytes = pl.dy_slot_bytes;
    cudaError_t e = cudaFuncSetAttribute(dw::conv3d_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (e != cudaSuccess) { set_error("conv3d_dw: cannot raise dynamic smem: %s", cudaGetErrorString(e)); return (int)e; }
    for (int gi = 0; gi < gin; ++gi)
        for (int go = 0; go < gout; ++go) {
            const int ci_off = gi * 32, co_off = go * 32;
            const int cin_g = gin == 1 ? desc->cin : (desc->cin - ci_off < 32 ? desc->cin - ci_off : 32);
            const int cout_g = gout == 1 ? desc->cout : (desc->cout - co_off < 32 ? desc->cout - co_off : 32);
            if (cin_g <= 0 || cout_g <= 0) continue;
            p.x_kc0 = gin == 1 ? 0 : gi * 4;
            p.dy_kc0 = gout == 1 ? 0 : go * 4;
            dw::conv3d_dw_kernel<<<pl.G * 3, dw::kThreads, pl.smem_bytes, st>>>(p);
            const int total = (pl.two_d ? 9 : 27) * cin_g * cout_g;
            dw::dw_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(ws, grad_w_packed, pl.G, pl.NM, pl.NCOPY, pl.R, pl.N, cin_g,
                                                                      cout_g, pl.cin_pad, pl.cout_pad, pl.nparts, desc->scale,
                                                                      pl.two_d, desc->cin, desc->cout, gin == 1 ? 0 : ci_off,
                                                                      gout == 1 ? 0 : co_off);
        }
    if (grad_bias != nullptr) {
        float* partial = ws + (int64_t)pl.G * 3 * pl.nacc * 128 * pl.N;
        const int planes = desc->n * desc->d;
        dw::split_colsum_kernel<<<planes, 256, 0, st>>>(p.dy, p.dy_part, pl.nparts, cout_pad / 8, pl.PP, partial);
        dw::colsum_finish_kernel<<<(desc->cout + 63) / 64, 64, 0, st>>>(partial, planes, cout_pad, desc->cout, grad_bias);
    }
    LF_RETURN_LAUNCH();
}
