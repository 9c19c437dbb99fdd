// K1/K2 — the two voxel resamplers of the LatentFusion hot path, sm_100a.
//
//   object->camera  (ObjectToCameraTransform, modules/geometry.py:669-690)
//   camera->object  (CameraToObjectTransform, modules/geometry.py:625-657)
//
// Both are a trilinear gather (F.grid_sample, padding_mode='border', align_corners=False;
// geometry.py:16-17) whose sampling grid is an analytic function of ~20 camera floats.  The
// reference materialises the grid ([N,S^3,3] + five intermediates) and N copies of the cube; here the
// grid is generated in registers from the per-camera constant block and the single cube is read
// through L2.  Layout is channels-last, so one voxel's C channels are one contiguous 4*C-byte run:
// a group of LPV = C/4 lanes reads each of the 8 corners as coalesced 128-bit loads and writes the
// result voxel as one contiguous run.
//
// HBM traffic (algorithmic): fwd 4*C*S^3*(B + N) bytes; bwd_cam 4*C*S^3*(N + B); bwd_vol same.
#include "common.cuh"
#include "tc_common.cuh"

namespace lf {

struct Samp {
    int x0, y0, z0;      // floor corner (west / north / top in ATen's naming)
    float wx1, wy1, wz1; // weight of the +1 corner along each axis (= frac)
    float wx0, wy0, wz0; // weight of the floor corner
    float mx, my, mz;    // d(ix)/d(grid coord): S/2 strictly inside, 0 where border-clamped
};

// ATen GridSampler.h: grid_sampler_unnormalize (align_corners=False) + clip_coordinates_set_grad.
__device__ __forceinline__ void unnorm_clip(float g, int S, float& ix, float& mult) {
    ix = ((g + 1.f) * (float)S - 1.f) / 2.f;
    const float mx = (float)(S - 1);
    if (ix <= 0.f) { ix = 0.f; mult = 0.f; }
    else if (ix >= mx) { ix = mx; mult = 0.f; }
    else { mult = (float)S / 2.f; }
}

__device__ __forceinline__ Samp make_samp(float gx, float gy, float gz, int S) {
    Samp s;
    float ix, iy, iz;
    unnorm_clip(gx, S, ix, s.mx);
    unnorm_clip(gy, S, iy, s.my);
    unnorm_clip(gz, S, iz, s.mz);
    float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    s.x0 = (int)fx; s.y0 = (int)fy; s.z0 = (int)fz;
    s.wx1 = ix - fx; s.wy1 = iy - fy; s.wz1 = iz - fz;
    s.wx0 = (fx + 1.f) - ix; s.wy0 = (fy + 1.f) - iy; s.wz0 = (fz + 1.f) - iz;
    return s;
}

// MODE 0: object->camera grid (geometry.py:469-531, :669-685).  voxel (k,j,i) = (depth, v, u).
// MODE 1: camera->object grid (geometry.py:599-611, :625-654).  voxel (k,j,i) = (z, y, x) lattice.
template <int MODE>
__device__ __forceinline__ void gen_grid(const float* __restrict__ cm, int S, int i, int j, int k,
                                         float& gx, float& gy, float& gz) {
    if (MODE == 0) {
        const float tu = linspace_at(0.f, 1.f, S, i);
        const float tv = linspace_at(0.f, 1.f, S, j);
        const float tz = linspace_at(0.f, 1.f, S, k);
        const float u = tu * cm[14] + cm[12];
        const float v = tv * cm[15] + cm[13];
        const float z = tz * cm[21] + cm[20];
        const float y = (v - cm[17]) / cm[19] * z;
        const float x = (u - cm[16]) / cm[18] * z;
        const float ox = cm[0] * x + cm[1] * y + cm[2] * z + cm[3];
        const float oy = cm[4] * x + cm[5] * y + cm[6] * z + cm[7];
        const float oz = cm[8] * x + cm[9] * y + cm[10] * z + cm[11];
        const float half = cm[22];   // cube_size / 2
        gx = ox / half; gy = oy / half; gz = oz / half;
    } else {
        const float hc = cm[30] * 0.5f;
        const float x = linspace_at(-hc, hc, S, i);
        const float y = linspace_at(-hc, hc, S, j);
        const float z = linspace_at(-hc, hc, S, k);
        const float cx = cm[0] * x + cm[1] * y + cm[2] * z + cm[3];
        const float cy = cm[4] * x + cm[5] * y + cm[6] * z + cm[7];
        const float cz = cm[8] * x + cm[9] * y + cm[10] * z + cm[11];
        const float p0 = cm[16] * cx + cm[17] * cy + cm[18] * cz + cm[19];
        const float p1 = cm[20] * cx + cm[21] * cy + cm[22] * cz + cm[23];
        const float p2 = cm[24] * cx + cm[25] * cy + cm[26] * cz + cm[27];
        const float px = p0 / p2, py = p1 / p2;
        gx = ((px - cm[12]) / cm[14]) * 2.f - 1.f;
        gy = ((py - cm[13]) / cm[15]) * 2.f - 1.f;
        gz = (p2 - cm[28]) / (cm[29] - cm[28]);   // NB: [0,1], no 2x-1 (reference quirk, kept)
    }
}

template <int VEC> struct Vec;
template <> struct Vec<4> {
    typedef float4 T;
    static __device__ __forceinline__ T load(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
    static __device__ __forceinline__ void store(float* p, T v) { __stcs(reinterpret_cast<float4*>(p), v); }
    static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ void fma(T& acc, float w, T v) {
        ffma2_bcast(w, v.x, v.y, acc.x, acc.y);
        ffma2_bcast(w, v.z, v.w, acc.z, acc.w);
    }
    static __device__ __forceinline__ float dot(T a, T b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
    static __device__ __forceinline__ T sub(T a, T b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
    static __device__ __forceinline__ T scale(T a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
    static __device__ __forceinline__ void atomic_add(float* p, T v) {
        atomicAdd(reinterpret_cast<float4*>(p), v);   // red.global.add.v4.f32 (sm_90+)
    }
};
template <> struct Vec<1> {
    typedef float T;
    static __device__ __forceinline__ T load(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ void store(float* p, T v) { __stcs(p, v); }
    static __device__ __forceinline__ T zero() { return 0.f; }
    static __device__ __forceinline__ void fma(T& acc, float w, T v) { acc += w * v; }
    static __device__ __forceinline__ float dot(T a, T b) { return a * b; }
    static __device__ __forceinline__ T sub(T a, T b) { return a - b; }
    static __device__ __forceinline__ T scale(T a, float s) { return a * s; }
    static __device__ __forceinline__ void atomic_add(float* p, T v) { atomicAdd(p, v); }
};

struct CornerOfs {
    unsigned o[8];  // element offsets (before channel) in ATen order: tnw tne tsw tse bnw bne bsw bse
    float w[8];     // (a single cube is < 2^31 elements, checked on the host, so 32-bit offsets suffice)
};

__device__ __forceinline__ CornerOfs corner_offsets(int x0, int y0, int z0, float wx1, float wy1, float wz1,
                                                    int S, int C) {
    CornerOfs c;
    const int sy = S * C, sz = S * sy;
    // a +1 corner that falls outside only happens at ix == S-1 where its weight is exactly 0: clamp it
    const int dx = (x0 + 1 < S) ? C : 0, dy = (y0 + 1 < S) ? sy : 0, dz = (z0 + 1 < S) ? sz : 0;
    const unsigned base = (unsigned)(z0 * sz + y0 * sy + x0 * C);
    const float wx0 = 1.f - wx1, wy0 = 1.f - wy1, wz0 = 1.f - wz1;   // == (floor+1) - ix exactly
    const float w00 = wy0 * wz0, w10 = wy1 * wz0, w01 = wy0 * wz1, w11 = wy1 * wz1;
    c.o[0] = base;                c.w[0] = wx0 * w00;
    c.o[1] = base + dx;           c.w[1] = wx1 * w00;
    c.o[2] = base + dy;           c.w[2] = wx0 * w10;
    c.o[3] = base + dy + dx;      c.w[3] = wx1 * w10;
    c.o[4] = base + dz;           c.w[4] = wx0 * w01;
    c.o[5] = base + dz + dx;      c.w[5] = wx1 * w01;
    c.o[6] = base + dz + dy;      c.w[6] = wx0 * w11;
    c.o[7] = base + dz + dy + dx; c.w[7] = wx1 * w11;
    return c;
}

__device__ __forceinline__ CornerOfs corner_offsets(const Samp& s, int S, int C) {
    return corner_offsets(s.x0, s.y0, s.z0, s.wx1, s.wy1, s.wz1, S, C);
}

// ------------------------------------------------------------------------------------------
// forward: one lane-group (LPV lanes) per output voxel
// ------------------------------------------------------------------------------------------
// Output voxels are walked brick by brick (BX x BY x BZ voxels per CTA) so that the 8-corner reuse
// between neighbouring output voxels is served by L1 instead of L2: a b^3 brick touches ~(b+1)^3 input
// voxels for 8*b^3 corner reads.
constexpr int BX = 8, BY = 8, BZ = 4;          // 256 voxels per CTA

struct BrickGrid {
    int nbx, nby, nbz;                          // bricks per axis
    __host__ __device__ int per_cam() const { return nbx * nby * nbz; }
};

__host__ __device__ inline BrickGrid brick_grid(int S, int bx, int by, int bz) {
    BrickGrid g;
    g.nbx = (S + bx - 1) / bx; g.nby = (S + by - 1) / by; g.nbz = (S + bz - 1) / bz;
    return g;
}

// Packed per-voxel sampling state exchanged between lanes: floor corner (10 bits per axis) + the
// three fractional weights.  (fx+1)-ix == 1-(ix-fx) exactly in fp32, so the floor-side weights are
// rebuilt on the receiving lane without changing a bit.
struct PackedSamp {
    int base;            // x0 | y0 << 10 | z0 << 20
    float wx1, wy1, wz1;
};

__device__ __forceinline__ PackedSamp pack_samp(const Samp& s) {
    PackedSamp p;
    p.base = s.x0 | (s.y0 << 10) | (s.z0 << 20);
    p.wx1 = s.wx1; p.wy1 = s.wy1; p.wz1 = s.wz1;
    return p;
}

__device__ __forceinline__ PackedSamp shfl_samp(const PackedSamp& p, int src) {
    PackedSamp r;
    r.base = __shfl_sync(0xffffffffu, p.base, src);
    r.wx1 = __shfl_sync(0xffffffffu, p.wx1, src);
    r.wy1 = __shfl_sync(0xffffffffu, p.wy1, src);
    r.wz1 = __shfl_sync(0xffffffffu, p.wz1, src);
    return r;
}

__device__ __forceinline__ CornerOfs corner_offsets(const PackedSamp& p, int S, int C) {
    return corner_offsets(p.base & 1023, (p.base >> 10) & 1023, (p.base >> 20) & 1023, p.wx1, p.wy1, p.wz1, S, C);
}

// Forward.  A warp owns 32 output voxels (an 8 x 4 slab of the CTA's 8x8x4 brick).  Phase 1: lane l
// generates the sample position of voxel l (the ~300-instruction camera chain runs once per 32 voxels,
// not once per lane-group).  Phase 2: the warp walks the 32 voxels 32/LPV at a time; the packed
// sampling state is broadcast with 4 shuffles and each LPV-lane group gathers its voxel's 8 corners
// with coalesced 128-bit loads and writes one contiguous 4*C-byte run.
template <int MODE, int VEC, bool ONE_CHUNK>
__global__ void __launch_bounds__(256)
resample_fwd_kernel(const float* __restrict__ vol, const float* __restrict__ cam, float* __restrict__ out,
                    int views_per_obj, int N, int C, int S, int lpv_log2) {
    const int64_t S3 = (int64_t)S * S * S;
    const BrickGrid bg = brick_grid(S, BX, BY, BZ);
    const int n = blockIdx.x / bg.per_cam();
    int b = blockIdx.x - n * bg.per_cam();
    const int bi = b % bg.nbx; b /= bg.nbx;
    const int bj = b % bg.nby; const int bk = b / bg.nby;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lpv = 1 << lpv_log2;
    const int sub = lane & (lpv - 1);
    const int grp = lane >> lpv_log2;          // voxel slot inside one gather step
    const int vps = 32 >> lpv_log2;            // voxels per gather step

    __shared__ float cm[LF_CAM_STRIDE];
    if (threadIdx.x < LF_CAM_STRIDE) cm[threadIdx.x] = cam[(int64_t)n * LF_CAM_STRIDE + threadIdx.x];
    __syncthreads();

    // phase 1: this lane's voxel
    const int lv = warp * 32 + lane;
    const int vi = bi * BX + (lv % BX), vj = bj * BY + (lv / BX) % BY, vk = bk * BZ + lv / (BX * BY);
    const bool inside = vi < S && vj < S && vk < S;
    PackedSamp mine;
    {
        float gx, gy, gz;
        gen_grid<MODE>(cm, S, min(vi, S - 1), min(vj, S - 1), min(vk, S - 1), gx, gy, gz);
        mine = pack_samp(make_samp(gx, gy, gz, S));
    }
    const unsigned inside_mask = __ballot_sync(0xffffffffu, inside);

    // phase 2
    const float* vb = vol + (int64_t)(MODE == 0 ? n / views_per_obj : n) * S3 * C;
    float* on = out + (int64_t)n * S3 * C;
    const int row0 = (bk * BZ + warp / 2) * S + bj * BY + (warp & 1) * 4;   // (k*S + j) of this warp's slab row 0
    for (int step = 0; step < lpv; ++step) {
        const int src = step * vps + grp;
        const PackedSamp ps = shfl_samp(mine, src);
        if (!((inside_mask >> src) & 1u)) continue;
        const CornerOfs co = corner_offsets(ps, S, C);
        const unsigned opos = (unsigned)(((row0 + (src >> 3)) * S + bi * BX + (src & 7)) * C);
        // one 64-bit base address per voxel; the other 7 corners are small non-negative deltas from it
        const unsigned d1 = co.o[1] - co.o[0], d2 = co.o[2] - co.o[0], d4 = co.o[4] - co.o[0];
        for (unsigned c = sub * VEC; c < (unsigned)C; c += lpv * VEC) {
            const float* p0 = vb + (co.o[0] + c);
            typename Vec<VEC>::T val[8];
            val[0] = Vec<VEC>::load(p0);
            val[1] = Vec<VEC>::load(p0 + d1);
            val[2] = Vec<VEC>::load(p0 + d2);
            val[3] = Vec<VEC>::load(p0 + (d2 + d1));
            val[4] = Vec<VEC>::load(p0 + d4);
            val[5] = Vec<VEC>::load(p0 + (d4 + d1));
            val[6] = Vec<VEC>::load(p0 + (d4 + d2));
            val[7] = Vec<VEC>::load(p0 + (d4 + d2 + d1));
            typename Vec<VEC>::T acc = Vec<VEC>::zero();
#pragma unroll
            for (int q = 0; q < 8; ++q) Vec<VEC>::fma(acc, co.w[q], val[q]);
            Vec<VEC>::store(on + (opos + c), acc);
            if (ONE_CHUNK) break;
        }
    }
}

// ------------------------------------------------------------------------------------------
// forward, depth-marching variant (C = 4 << LPVL, i.e. one 128-bit chunk per lane)
// ------------------------------------------------------------------------------------------
// The brick kernel above re-reads all 8 corners of every output voxel through L1 (9 L1 wavefronts and ~40
// warp instructions per voxel: co-limited by the SM's LSU and issue rates at ~37 % of the HBM roofline).  Here
// a lane group owns one (i, j) output column and walks it in depth.  Consecutive samples of a column are
// ~half a source voxel apart, so most of the 8 corners of step k are corners of step k-1 as well.  They
// are kept in registers, addressed by the PARITY of their source coordinate: the 2x2x2 cell
// {x0, x0+1} x {y0, y0+1} x {z0, z0+1} holds exactly one voxel of each parity class (px, py, pz), so slot
// (px, py, pz) always has a unique owner, nothing ever moves between registers, and a slot is re-loaded
// only when the voxel it must hold changes (measured 3.7 loads per step instead of 8).  The per-voxel sampling state (8 slot offsets + 8 slot weights) is computed once by one lane —
// lane (g, s) of a warp prepares step s of column g for the next LPV steps — and handed over through shared
// memory, so the lanes of a group do not repeat the index/weight arithmetic; the per-column part of the camera
// chain (two IEEE divisions) is hoisted out of the depth loop.  ~26 warp instructions per voxel instead of 40.
// Measured at config B (ncu, profiles/): 102 us vs 118 us for the brick kernel; the remaining limiter is the L2
// round trip of the ~3.7 new lines of a step (the FMAs of a step wait for its own loads; 24 warps per SM at 80
// registers).  Variants measured and dropped: prefetch.global.L1 one round ahead (CCTL.PF1: no gain, lower L1 hit
// rate), 4 CTAs/SM at 64 registers (150 us: the larger shared-memory carve-out costs more L1 than the extra
// warps give), 2x2 column patches per warp, block-level lockstep (both neutral or worse).  Round 2 (profiles/r02_*):
// 256-bit accesses (LDG.E.ENL2.256, W = 2 below: half the warp instructions per byte, 118 registers, 2 CTAs/SM) 124 us;
// a one-step software pipeline (the slots that change are loaded into a second register set during the previous step's
// arithmetic, records prepared one round ahead) cuts the long-scoreboard stall from 6.4 to 2.1 per issue but needs 122
// registers and +47 % instructions (the predicated register moves that commit the new lines): 118 us.  The kernel sits at
// ~50 % issue, ~56 % L1 data pipe and ~57 % long-scoreboard samples at once: every variant that relieves one of the three
// loads another, so the 128-bit / 3 CTAs per SM form stays the default.
// 128-bit read-only load at base + off (float4 units): one IMAD.WIDE for the address instead of a 64-bit add chain
__device__ __forceinline__ float4 ldg_f4_at(const float4* base, uint32_t off) {
    unsigned long long addr;
    asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(addr) : "r"(off), "l"(base));
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr));
    return v;
}

// 256-bit variant (sm_100: LDG.E.ENL2.256): 8 channels per lane, so a voxel of C channels is C/8 lanes and one warp
// instruction moves twice the voxels — half the LSU / address / predicate instructions per byte of the 128-bit form
__device__ __forceinline__ void ldg_f8_at(const float4* base, uint32_t off, float4& a, float4& b) {
    unsigned long long addr;
    asm("mad.wide.u32 %0, %1, 32, %2;" : "=l"(addr) : "r"(off), "l"(base));
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(addr));
}
__device__ __forceinline__ void stcs_f8(float4* p, const float4& a, const float4& b) {
    asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

// W = float4s per lane (1: 128-bit accesses, C = 4*LPV; 2: 256-bit accesses, C = 8*LPV)
// SPLIT (object->camera, W = 1): `out` is the library's split-planar activation layout ([hi|lo][N*S planes][C/8][S+2][S+2][8]
// bf16, what the depth-batched convolution stages with bulk TMA) instead of dense fp32: the first camera-block
// convolution then needs no packing pass (268 MB read + 285 MB written per pose-loop iteration at config B).  Same bytes
// written as fp32; each lane stores its 4 channels as 8 bytes of the hi part and 8 bytes of the lo part; the zero halo is
// laid down by split_halo_zero_kernel.  Measured at config B: 182 us (with the halo kernel) against 100 us for the dense
// form + 100 us for lf_split_pack — the 8-byte pieces land in 8 planes per warp step, twice the store wavefronts of one
// 512-byte run, on a kernel whose L1 data pipe is already a co-limiter — so the Photographer only uses it on request
// (LFB200_O2C_SPLIT=1); it halves the HBM traffic of the pair but not its time.
template <int MODE, int LPVL, int MINB, int W, bool SPLIT = false>
__global__ void __launch_bounds__(256, MINB)
resample_march_kernel(const float* __restrict__ vol, const float* __restrict__ cam, float* __restrict__ out,
                      int views_per_obj, int N, int S, int KC) {
    constexpr int LPV = 1 << LPVL, G = 32 / LPV, C = 4 * W * LPV;
    constexpr int TI = 2 * G, TJ = 4;
    // record buffer of one warp and one round: 4 components (offsets 0-3, offsets 4-7, weights 0-3, weights 4-7) x
    // one 16-byte slot per (group, step); a group's slots are followed by one pad slot so that the 4 groups of a
    // warp read different banks, and a warp-wide store (lane = slot) is conflict-free.
    constexpr int SLOTS = G * (LPV + 1);
    const int nti = (S + TI - 1) / TI, ntj = (S + TJ - 1) / TJ, nkc = (S + KC - 1) / KC;
    int b = blockIdx.x;
    const int ti = b % nti; b /= nti;
    const int tj = b % ntj; b /= ntj;
    const int kc = b % nkc; const int n = b / nkc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & (LPV - 1), g = lane >> LPVL;
    const int i = ti * TI + (warp & 1) * G + g;
    const int j = tj * TJ + (warp >> 1);
    const bool col_ok = i < S && j < S;
    const int ic = min(i, S - 1), jc = min(j, S - 1);
    const int k0 = kc * KC, k1 = min(S, k0 + KC);

    __shared__ float cm[LF_CAM_STRIDE];
    __shared__ uint4 rec[8][4][SLOTS];
    if (threadIdx.x < LF_CAM_STRIDE) cm[threadIdx.x] = cam[(int64_t)n * LF_CAM_STRIDE + threadIdx.x];
    __syncthreads();

    // per-column part of the object->camera chain (same operations, in the same order, as gen_grid<0>)
    float ax = 0.f, by = 0.f;
    if (MODE == 0) {
        const float u = linspace_at(0.f, 1.f, S, ic) * cm[14] + cm[12];
        const float v = linspace_at(0.f, 1.f, S, jc) * cm[15] + cm[13];
        ax = (u - cm[16]) / cm[18];
        by = (v - cm[17]) / cm[19];
    }

    const int64_t S3 = (int64_t)S * S * S;
    const float4* vcube = reinterpret_cast<const float4*>(vol + (int64_t)(MODE == 0 ? n / views_per_obj : n) * S3 * C);
    const float4* vb = vcube + sub * W;
    float4* op = reinterpret_cast<float4*>(out + ((((int64_t)n * S + k0) * S + jc) * S + ic) * C) + sub * W;
    const int64_t ostep = (int64_t)S * S * LPV * W;        // float4 units per depth step
    const int slot = g * (LPV + 1) + sub;
    // split-planar output: channel chunk sub/2 of plane (n, k), padded position (jc+1, ic+1), half (sub&1) of its 16 bytes
    uint16_t* sp = nullptr;
    int64_t sp_step = 0, sp_part = 0;
    if (SPLIT) {
        constexpr int KCH = C / 8;
        const int Wp = S + 2;
        const int64_t PP = (int64_t)(S + 2) * Wp;
        sp_part = (int64_t)N * S * C * PP;
        sp_step = (int64_t)KCH * PP * 8;
        sp = reinterpret_cast<uint16_t*>(out) + ((((int64_t)n * S + k0) * KCH + (sub >> 1)) * PP + (int64_t)(jc + 1) * Wp + (ic + 1)) * 8
             + (sub & 1) * 4;
    }

    // ---- phase 1: lane (g, sub) prepares depth step kr + sub of column g
    auto prepare = [&](int kr) {
        const int k = min(kr + sub, S - 1);
        float gx, gy, gz;
        if (MODE == 0) {
            const float z = linspace_at(0.f, 1.f, S, k) * cm[21] + cm[20];
            const float y = by * z;
            const float x = ax * z;
            const float ox = cm[0] * x + cm[1] * y + cm[2] * z + cm[3];
            const float oy = cm[4] * x + cm[5] * y + cm[6] * z + cm[7];
            const float oz = cm[8] * x + cm[9] * y + cm[10] * z + cm[11];
            const float half = cm[22];
            gx = ox / half; gy = oy / half; gz = oz / half;
        } else {
            gen_grid<1>(cm, S, ic, jc, k, gx, gy, gz);
        }
        const Samp sp = make_samp(gx, gy, gz, S);
        // slot p of an axis holds the cell's voxel whose coordinate has parity p; a +1 corner outside the
        // volume only occurs with weight exactly 0 and is clamped onto the last voxel
        const int x1 = min(sp.x0 + 1, S - 1), y1 = min(sp.y0 + 1, S - 1), z1 = min(sp.z0 + 1, S - 1);
        const float wx0 = 1.f - sp.wx1, wy0 = 1.f - sp.wy1, wz0 = 1.f - sp.wz1;
        const bool ox_ = sp.x0 & 1, oy_ = sp.y0 & 1, oz_ = sp.z0 & 1;
        const int X[2] = {ox_ ? x1 : sp.x0, ox_ ? sp.x0 : x1};
        const int Y[2] = {oy_ ? y1 : sp.y0, oy_ ? sp.y0 : y1};
        const int Z[2] = {oz_ ? z1 : sp.z0, oz_ ? sp.z0 : z1};
        const float WX[2] = {ox_ ? sp.wx1 : wx0, ox_ ? wx0 : sp.wx1};
        const float WY[2] = {oy_ ? sp.wy1 : wy0, oy_ ? wy0 : sp.wy1};
        const float WZ[2] = {oz_ ? sp.wz1 : wz0, oz_ ? wz0 : sp.wz1};
        uint32_t off[8];
        float w[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int px = q & 1, py = (q >> 1) & 1, pz = q >> 2;
            off[q] = (uint32_t)((Z[pz] * S + Y[py]) * S + X[px]) * LPV;       // units of one lane's W float4s
            w[q] = WX[px] * (WY[py] * WZ[pz]);
        }
        rec[warp][0][slot] = make_uint4(off[0], off[1], off[2], off[3]);
        rec[warp][1][slot] = make_uint4(off[4], off[5], off[6], off[7]);
        rec[warp][2][slot] = make_uint4(__float_as_uint(w[0]), __float_as_uint(w[1]), __float_as_uint(w[2]), __float_as_uint(w[3]));
        rec[warp][3][slot] = make_uint4(__float_as_uint(w[4]), __float_as_uint(w[5]), __float_as_uint(w[6]), __float_as_uint(w[7]));
    };

    uint32_t held[8];
    float4 val[8][W];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        held[q] = 0xffffffffu;
#pragma unroll
        for (int h = 0; h < W; ++h) val[q][h] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    for (int kr = k0; kr < k1; kr += LPV) {
        prepare(kr);
        __syncwarp();
        // ---- phase 2: the group walks its column through the records of this round
        const int nsteps = min(LPV, k1 - kr);
        const uint4* r0 = &rec[warp][0][g * (LPV + 1)];
#pragma unroll 2
        for (int s = 0; s < nsteps; ++s) {
            const uint4 o0 = r0[s], o1 = r0[SLOTS + s], w0 = r0[2 * SLOTS + s], w1 = r0[3 * SLOTS + s];
            const uint32_t off[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
            const float w[8] = {__uint_as_float(w0.x), __uint_as_float(w0.y), __uint_as_float(w0.z), __uint_as_float(w0.w),
                                __uint_as_float(w1.x), __uint_as_float(w1.y), __uint_as_float(w1.z), __uint_as_float(w1.w)};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (col_ok && off[q] != held[q]) {
                    if (W == 1) val[q][0] = ldg_f4_at(vb, off[q]);
                    else ldg_f8_at(vb, off[q], val[q][0], val[q][W - 1]);
                }
                held[q] = off[q];          // (a slot is current after every step: no conditional bookkeeping)
            }
            float4 acc[W];
#pragma unroll
            for (int h = 0; h < W; ++h) acc[h] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
#pragma unroll
                for (int h = 0; h < W; ++h) Vec<4>::fma(acc[h], w[q], val[q][h]);
            }
            if (col_ok) {
                if (SPLIT) {
                    uint32_t h0, l0, h1, l1;
                    tcx::split_bf16x2(acc[0].x, acc[0].y, h0, l0);
                    tcx::split_bf16x2(acc[0].z, acc[0].w, h1, l1);
                    *reinterpret_cast<uint2*>(sp) = make_uint2(h0, h1);
                    *reinterpret_cast<uint2*>(sp + sp_part) = make_uint2(l0, l1);
                } else if (W == 1) __stcs(op, acc[0]);
                else stcs_f8(op, acc[0], acc[W - 1]);
            }
            op += ostep;
            if (SPLIT) sp += sp_step;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// backward w.r.t. the volume: trilinear splat with fp32 vector reductions into L2
// ------------------------------------------------------------------------------------------
template <int MODE, int VEC>
__global__ void __launch_bounds__(256)
resample_bwd_vol_kernel(const float* __restrict__ gout, const float* __restrict__ cam, float* __restrict__ gvol,
                        int views_per_obj, int N, int C, int S, int lpv_log2) {
    const int64_t S3 = (int64_t)S * S * S;
    const int64_t total = (int64_t)N * S3;
    const int lpv = 1 << lpv_log2;
    const int64_t gthread = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int sub = (int)(gthread & (lpv - 1));
    const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) >> lpv_log2;
    for (int64_t v = gthread >> lpv_log2; v < total; v += ngroups) {
        const int n = (int)(v / S3);
        const int p = (int)(v - (int64_t)n * S3);
        const int i = p % S, j = (p / S) % S, k = p / (S * S);
        float gx, gy, gz;
        gen_grid<MODE>(cam + (int64_t)n * LF_CAM_STRIDE, S, i, j, k, gx, gy, gz);
        const Samp s = make_samp(gx, gy, gz, S);
        const CornerOfs co = corner_offsets(s, S, C);
        float* vb = gvol + (int64_t)(MODE == 0 ? n / views_per_obj : n) * S3 * C;
        const float* gb = gout + v * C;
        for (int c = sub * VEC; c < C; c += lpv * VEC) {
            const typename Vec<VEC>::T g = Vec<VEC>::load(gb + c);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (co.w[q] != 0.f) Vec<VEC>::atomic_add(vb + co.o[q] + c, Vec<VEC>::scale(g, co.w[q]));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward w.r.t. the camera block (object->camera only; the pose loop's gradient).
// Stage 1: each block owns a contiguous voxel chunk of ONE camera, accumulates the 17 partials in
// registers -> warp shuffle -> shared -> workspace[n][block][17].  Stage 2: fixed-order sum per
// camera (deterministic, no atomics).
// ------------------------------------------------------------------------------------------
constexpr int kCamGradTerms = 17;          // M[12], vp x0,y0,w,h, znear
constexpr int CBX = 16, CBY = 8, CBZ = 8;   // bwd_cam brick: 1024 voxels per block


template <int VEC>
__global__ void __launch_bounds__(256, 3)
resample_o2c_bwd_cam_kernel(const float* __restrict__ gout, const float* __restrict__ vol,
                            const float* __restrict__ cam, float* __restrict__ ws,
                            int views_per_obj, int N, int C, int S, int lpv_log2, int blocks_per_cam) {
    const int S3 = S * S * S;
    const int n = blockIdx.x / blocks_per_cam;
    const int b = blockIdx.x - n * blocks_per_cam;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lpv = 1 << lpv_log2;
    const int sub = lane & (lpv - 1);
    const int grp = lane >> lpv_log2;
    const int vps = 32 >> lpv_log2;

    __shared__ float cm[LF_CAM_STRIDE];
    if (threadIdx.x < LF_CAM_STRIDE) cm[threadIdx.x] = cam[(int64_t)n * LF_CAM_STRIDE + threadIdx.x];
    __syncthreads();

    float acc[kCamGradTerms];
#pragma unroll
    for (int t = 0; t < kCamGradTerms; ++t) acc[t] = 0.f;

    const float* vb = vol + (int64_t)(n / views_per_obj) * S3 * C;
    const float* gn = gout + (int64_t)n * S3 * C;
    const BrickGrid bg = brick_grid(S, CBX, CBY, CBZ);
    int bb_ = b;
    const int bi = bb_ % bg.nbx; bb_ /= bg.nbx;
    const int bj = bb_ % bg.nby; const int bk = bb_ / bg.nby;

    // the 16x8x8 brick is walked as 4 sub-bricks of 8x8x4; a warp owns an 8x4 slab of each
    for (int sb = 0; sb < 4; ++sb) {
        const int lv = warp * 32 + lane;
        const int vi = bi * CBX + (sb & 1) * 8 + (lv % 8);
        const int vj = bj * CBY + (lv / 8) % 8;
        const int vk = bk * CBZ + (sb >> 1) * 4 + lv / 64;
        const bool inside = vi < S && vj < S && vk < S;
        const int i = min(vi, S - 1), j = min(vj, S - 1), k = min(vk, S - 1);
        // --- phase 1: forward recompute of this lane's grid point (same op order as gen_grid<0>),
        //     keeping the intermediates the chain rule needs
        const float tu = linspace_at(0.f, 1.f, S, i);
        const float tv = linspace_at(0.f, 1.f, S, j);
        const float tz = linspace_at(0.f, 1.f, S, k);
        const float u = tu * cm[14] + cm[12];
        const float v = tv * cm[15] + cm[13];
        const float z = tz * cm[21] + cm[20];
        const float a = (u - cm[16]) / cm[18];
        const float bb = (v - cm[17]) / cm[19];
        const float x = a * z, y = bb * z;
        const float half = cm[22];
        const float gx = (cm[0] * x + cm[1] * y + cm[2] * z + cm[3]) / half;
        const float gy = (cm[4] * x + cm[5] * y + cm[6] * z + cm[7]) / half;
        const float gz = (cm[8] * x + cm[9] * y + cm[10] * z + cm[11]) / half;
        const Samp s = make_samp(gx, gy, gz, S);
        const PackedSamp mine = pack_samp(s);
        const int my_pos = inside ? (k * S + j) * S + i : -1;

        // --- phase 2: gather + contraction with grad_out, 32/LPV voxels per step
        float mdx = 0.f, mdy = 0.f, mdz = 0.f;      // d(sum_c go*sample)/d(ix,iy,iz) of MY voxel
        for (int step = 0; step < lpv; ++step) {
            const int src = step * vps + grp;
            const PackedSamp ps = shfl_samp(mine, src);
            const int pos = __shfl_sync(0xffffffffu, my_pos, src);
            float dx = 0.f, dy = 0.f, dz = 0.f;
            if (pos >= 0) {
                const CornerOfs co = corner_offsets(ps, S, C);
                const float wx0 = 1.f - ps.wx1, wy0 = 1.f - ps.wy1, wz0 = 1.f - ps.wz1;
                const float* gb = gn + (int64_t)pos * C;
                for (int c = sub * VEC; c < C; c += lpv * VEC) {
                    typename Vec<VEC>::T val[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) val[q] = Vec<VEC>::load(vb + co.o[q] + c);
                    const typename Vec<VEC>::T g = Vec<VEC>::load(gb + c);
                    // corner order: q = zbit*4 + ybit*2 + xbit
                    float d[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) d[q] = Vec<VEC>::dot(g, val[q]);
                    dx += (d[1] - d[0]) * wy0 * wz0 + (d[3] - d[2]) * ps.wy1 * wz0
                        + (d[5] - d[4]) * wy0 * ps.wz1 + (d[7] - d[6]) * ps.wy1 * ps.wz1;
                    dy += (d[2] - d[0]) * wx0 * wz0 + (d[3] - d[1]) * ps.wx1 * wz0
                        + (d[6] - d[4]) * wx0 * ps.wz1 + (d[7] - d[5]) * ps.wx1 * ps.wz1;
                    dz += (d[4] - d[0]) * wx0 * wy0 + (d[5] - d[1]) * ps.wx1 * wy0
                        + (d[6] - d[2]) * wx0 * ps.wy1 + (d[7] - d[3]) * ps.wx1 * ps.wy1;
                }
            }
            for (int o = lpv >> 1; o > 0; o >>= 1) {
                dx += __shfl_xor_sync(0xffffffffu, dx, o);
                dy += __shfl_xor_sync(0xffffffffu, dy, o);
                dz += __shfl_xor_sync(0xffffffffu, dz, o);
            }
            // hand the result back to the lane that owns this voxel (lane == step*vps + group index)
            const int from = (lane - step * vps) << lpv_log2;     // leader lane of my group, if mine is in this step
            const float rx = __shfl_sync(0xffffffffu, dx, from & 31);
            const float ry = __shfl_sync(0xffffffffu, dy, from & 31);
            const float rz = __shfl_sync(0xffffffffu, dz, from & 31);
            if (lane >= step * vps && lane < (step + 1) * vps) { mdx = rx; mdy = ry; mdz = rz; }
        }

        // --- phase 3: chain rule for my voxel: grid -> object coords -> (M, x, y, z) -> (viewport, znear)
        if (inside) {
            const float gox = mdx * s.mx / half, goy = mdy * s.my / half, goz = mdz * s.mz / half;
            acc[0] += gox * x; acc[1] += gox * y; acc[2] += gox * z; acc[3] += gox;
            acc[4] += goy * x; acc[5] += goy * y; acc[6] += goy * z; acc[7] += goy;
            acc[8] += goz * x; acc[9] += goz * y; acc[10] += goz * z; acc[11] += goz;
            const float lx = cm[0] * gox + cm[4] * goy + cm[8] * goz;
            const float ly = cm[1] * gox + cm[5] * goy + cm[9] * goz;
            const float lz = cm[2] * gox + cm[6] * goy + cm[10] * goz;
            const float lu = lx * z / cm[18];     // dL/du
            const float lv2 = ly * z / cm[19];    // dL/dv
            acc[12] += lu; acc[14] += lu * tu;
            acc[13] += lv2; acc[15] += lv2 * tv;
            acc[16] += lz + lx * a + ly * bb;     // dL/dznear  (z = tz*z_span + znear)
        }
    }

    // block reduction
    __shared__ float red[8][kCamGradTerms];
#pragma unroll
    for (int t = 0; t < kCamGradTerms; ++t) {
        const float r = warp_sum(acc[t]);
        if (lane == 0) red[warp][t] = r;
    }
    __syncthreads();
    if (threadIdx.x < kCamGradTerms) {
        float r = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += red[w][threadIdx.x];
        ws[((int64_t)n * blocks_per_cam + b) * kCamGradTerms + threadIdx.x] = r;
    }
}

// ------------------------------------------------------------------------------------------
// backward w.r.t. the camera block, depth-marching variant (C = 4 << LPVL)
// ------------------------------------------------------------------------------------------
// The brick kernel above spends ~84 warp instructions per voxel (ncu, profiles/r01b): every lane of a group
// repeats the corner index arithmetic, all 8 corners are re-read through L1, and every voxel pays 17 shuffles
// (state broadcast, 3-value tree over the group's lanes, hand-back to the owning lane).  This variant walks
// (i, j) columns in depth exactly like resample_march_kernel — parity-addressed register slots for the 8
// corners (~3.7 loads per step), per-step records prepared by one lane and handed over through shared memory —
// and removes the per-voxel cross-lane traffic altogether: along one column a = (u-cx)/fx, b = (v-cy)/fy, tu, tv
// are constants and x = a z, y = b z, so all 17 gradient terms are linear in six per-lane column sums
//     sum g_ox, sum g_ox z, sum g_oy, sum g_oy z, sum g_oz, sum g_oz z       (g_o* = d/d(ix,iy,iz) * mult / half)
// of the lane's OWN 4-channel partial dot products.  The lanes of a group never exchange anything until the single
// block reduction at the end of the chunk.  The three coordinate derivatives come from one hierarchical pass over
// the 8 per-corner dots d[q] = <g, corner q>: x-differences and x-interpolants first, then y, then z (27
// operations instead of 36).  grad_out (read once, from HBM) is prefetched D depth steps ahead into a register
// ring with L1::no_allocate so it does not evict the cube's lines; the corner slots that change at step s+1 are
// re-loaded right after the dot products of step s, so they are in flight during its remaining arithmetic.
// Measured at config B (8 cameras, 64^3 x 32 channels; tools/kbench.py --only bwdcam, ncu in profiles/r02c_*):
// 86 M warp instructions instead of 176 M, 278 -> 133 us (2.27 TB/s of algorithmic bytes, 34 % of the HBM roofline).
// The kernel is occupancy-bound (long-scoreboard stalls at 16-20 warps per SM): 96 registers and 4-warp CTAs (20
// warps) beat 128 registers and 8-warp CTAs (16 warps, 151 us); forcing 80 registers spills the corner slots.
__device__ __forceinline__ float4 ldg_f4_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// <a, b> with two packed operations and one add
__device__ __forceinline__ float dot4_x2(const float4& a, const float4& b) {
    unsigned long long a0, a1, b0, b1, t;
    a0 = ((unsigned long long)__float_as_uint(a.y) << 32) | __float_as_uint(a.x);
    a1 = ((unsigned long long)__float_as_uint(a.w) << 32) | __float_as_uint(a.z);
    b0 = ((unsigned long long)__float_as_uint(b.y) << 32) | __float_as_uint(b.x);
    b1 = ((unsigned long long)__float_as_uint(b.w) << 32) | __float_as_uint(b.z);
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(a0), "l"(b0));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(t) : "l"(a1), "l"(b1));
    return __uint_as_float((unsigned)t) + __uint_as_float((unsigned)(t >> 32));
}

constexpr int kBwdCamMaxChunks = 4;          // depth chunks per column (bounds the workspace, see lf_resample_o2c_bwd_cam_ws)

template <int LPVL, int NW, int MINB, int D>
__global__ void __launch_bounds__(NW * 32, MINB)
resample_o2c_bwd_cam_march_kernel(const float* __restrict__ gout, const float* __restrict__ vol,
                                  const float* __restrict__ cam, float* __restrict__ ws,
                                  int views_per_obj, int N, int S, int KC) {
    constexpr int LPV = 1 << LPVL, G = 32 / LPV, C = 4 * LPV;
    constexpr int TI = 2 * G, TJ = NW / 2;
    constexpr int SLOTS = G * (LPV + 1);
    
    const int nti = (S + TI - 1) / TI, ntj = (S + TJ - 1) / TJ, nkc = (S + KC - 1) / KC;
    int b = blockIdx.x;
    const int ti = b % nti; b /= nti;
    const int tj = b % ntj; b /= ntj;
    const int kc = b % nkc; const int n = b / nkc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & (LPV - 1), g = lane >> LPVL;
    const int i = ti * TI + (warp & 1) * G + g;
    const int j = tj * TJ + (warp >> 1);
    const bool col_ok = i < S && j < S;
    const int ic = min(i, S - 1), jc = min(j, S - 1);
    const int k0 = kc * KC, k1 = min(S, k0 + KC);

    __shared__ float cm[LF_CAM_STRIDE];
    // record of one (group, step): offsets 0-3 | offsets 4-7 | WX0 WX1 WY0 WY1 | WZ0 WZ1 z - | mx my mz -
    __shared__ uint4 rec[NW][5][SLOTS];
    __shared__ float red[NW][kCamGradTerms];
    if (threadIdx.x < LF_CAM_STRIDE) cm[threadIdx.x] = cam[(int64_t)n * LF_CAM_STRIDE + threadIdx.x];
    __syncthreads();

    // per-column part of the chain (same operations, in the same order, as gen_grid<0>)
    const float tu = linspace_at(0.f, 1.f, S, ic), tv = linspace_at(0.f, 1.f, S, jc);
    const float ax = (tu * cm[14] + cm[12] - cm[16]) / cm[18];
    const float by = (tv * cm[15] + cm[13] - cm[17]) / cm[19];
    const float half = cm[22];
    const float mh = ((float)S / 2.f) / half;          // border multiplier (S/2 inside, 0 where clamped) over half

    const int64_t S3 = (int64_t)S * S * S;
    const float4* vb = reinterpret_cast<const float4*>(vol + (int64_t)(n / views_per_obj) * S3 * C) + sub;
    const float4* gp = reinterpret_cast<const float4*>(gout + ((((int64_t)n * S + k0) * S + jc) * S + ic) * C) + sub;
    const int64_t ostep = (int64_t)S * S * LPV;        // float4 units per depth step
    const int slot = g * (LPV + 1) + sub;

    // lane (g, sub) prepares depth step kr + sub of column g.  Steps past the end of the chunk get a valid (clamped)
    // record with zero multipliers, so the walk below is straight-line code for all LPV steps of a round.
    auto prepare = [&](int kr) {
        const int k = min(kr + sub, S - 1);
        const float z = linspace_at(0.f, 1.f, S, k) * cm[21] + cm[20];
        const float y = by * z;
        const float x = ax * z;
        const float gx = (cm[0] * x + cm[1] * y + cm[2] * z + cm[3]) / half;
        const float gy = (cm[4] * x + cm[5] * y + cm[6] * z + cm[7]) / half;
        const float gz = (cm[8] * x + cm[9] * y + cm[10] * z + cm[11]) / half;
        const Samp sp = make_samp(gx, gy, gz, S);
        const int x1 = min(sp.x0 + 1, S - 1), y1 = min(sp.y0 + 1, S - 1), z1 = min(sp.z0 + 1, S - 1);
        const float wx0 = 1.f - sp.wx1, wy0 = 1.f - sp.wy1, wz0 = 1.f - sp.wz1;
        const bool ox_ = sp.x0 & 1, oy_ = sp.y0 & 1, oz_ = sp.z0 & 1;
        const int X[2] = {ox_ ? x1 : sp.x0, ox_ ? sp.x0 : x1};
        const int Y[2] = {oy_ ? y1 : sp.y0, oy_ ? sp.y0 : y1};
        const int Z[2] = {oz_ ? z1 : sp.z0, oz_ ? sp.z0 : z1};
        uint32_t off[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) off[q] = (uint32_t)((Z[q >> 2] * S + Y[(q >> 1) & 1]) * S + X[q & 1]) * LPV;
        rec[warp][0][slot] = make_uint4(off[0], off[1], off[2], off[3]);
        rec[warp][1][slot] = make_uint4(off[4], off[5], off[6], off[7]);
        rec[warp][2][slot] = make_uint4(__float_as_uint(ox_ ? sp.wx1 : wx0), __float_as_uint(ox_ ? wx0 : sp.wx1),
                                        __float_as_uint(oy_ ? sp.wy1 : wy0), __float_as_uint(oy_ ? wy0 : sp.wy1));
        rec[warp][3][slot] = make_uint4(__float_as_uint(oz_ ? sp.wz1 : wz0), __float_as_uint(oz_ ? wz0 : sp.wz1),
                                        __float_as_uint(z), 0u);
        // d(weight of slot 1)/d(ix) = +1 when slot 1 holds the +1 corner (x0 even), -1 otherwise; times the border
        // multiplier over half (sp.m* is S/2 strictly inside, 0 where clamped)
        const float live = (kr + sub < k1) ? mh : 0.f;
        rec[warp][4][slot] = make_uint4(__float_as_uint(sp.mx != 0.f ? (ox_ ? -live : live) : 0.f),
                                        __float_as_uint(sp.my != 0.f ? (oy_ ? -live : live) : 0.f),
                                        __float_as_uint(sp.mz != 0.f ? (oz_ ? -live : live) : 0.f), 0u);
    };

    uint32_t held[8];
    float4 val[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { held[q] = 0xffffffffu; val[q] = make_float4(0.f, 0.f, 0.f, 0.f); }
    float4 gring[D];
#pragma unroll
    for (int d = 0; d < D; ++d) gring[d] = (k0 + d < k1) ? ldg_f4_stream(gp + d * ostep) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* gnext = gp + D * ostep;              // grad_out of step (current + D)
    float A0 = 0.f, A1 = 0.f, B0 = 0.f, B1 = 0.f, C0 = 0.f, C1 = 0.f;

    const uint4* r0 = &rec[warp][0][g * (LPV + 1)];
    // the corner slots that change at step s are re-loaded (after the dot products of step s-1 have consumed them)
    auto fetch = [&](int s) {
        const uint4 o0 = r0[s], o1 = r0[SLOTS + s];
        const uint32_t off[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (col_ok && off[q] != held[q]) val[q] = ldg_f4_at(vb, off[q]);
            held[q] = off[q];
        }
    };

    for (int kr = k0; kr < k1; kr += LPV) {
        prepare(kr);
        __syncwarp();
        fetch(0);
#pragma unroll
        for (int s = 0; s < LPV; ++s) {
            float d[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) d[q] = dot4_x2(gring[s % D], val[q]);
            if (kr + s + D < k1) gring[s % D] = ldg_f4_stream(gnext);
            gnext += ostep;
            if (s + 1 < LPV) fetch(s + 1);              // in flight during the arithmetic below
            const uint4 wa = r0[2 * SLOTS + s], wb = r0[3 * SLOTS + s], wm = r0[4 * SLOTS + s];
            const float WX0 = __uint_as_float(wa.x), WX1 = __uint_as_float(wa.y);
            const float WY0 = __uint_as_float(wa.z), WY1 = __uint_as_float(wa.w);
            const float WZ0 = __uint_as_float(wb.x), WZ1 = __uint_as_float(wb.y), z = __uint_as_float(wb.z);
            // slot q = pz*4 + py*2 + px: differences and interpolants along x ...
            const float e0 = d[1] - d[0], e1 = d[3] - d[2], e2 = d[5] - d[4], e3 = d[7] - d[6];
            const float m0 = WX0 * d[0] + WX1 * d[1], m1 = WX0 * d[2] + WX1 * d[3];
            const float m2 = WX0 * d[4] + WX1 * d[5], m3 = WX0 * d[6] + WX1 * d[7];
            // ... then y, then z
            const float ex = WZ0 * (WY0 * e0 + WY1 * e1) + WZ1 * (WY0 * e2 + WY1 * e3);
            const float ey = WZ0 * (m1 - m0) + WZ1 * (m3 - m2);
            const float ez = (WY0 * m2 + WY1 * m3) - (WY0 * m0 + WY1 * m1);
            const float px = ex * __uint_as_float(wm.x), py = ey * __uint_as_float(wm.y), pz = ez * __uint_as_float(wm.z);
            A0 += px; A1 += px * z;
            B0 += py; B1 += py * z;
            C0 += pz; C1 += pz * z;
        }
        __syncwarp();
    }

    // six column sums -> the 17 terms (x = ax z, y = by z along the column)
    float acc[kCamGradTerms];
    acc[0] = ax * A1; acc[1] = by * A1; acc[2] = A1; acc[3] = A0;
    acc[4] = ax * B1; acc[5] = by * B1; acc[6] = B1; acc[7] = B0;
    acc[8] = ax * C1; acc[9] = by * C1; acc[10] = C1; acc[11] = C0;
    const float lx1 = cm[0] * A1 + cm[4] * B1 + cm[8] * C1, lx0 = cm[0] * A0 + cm[4] * B0 + cm[8] * C0;
    const float ly1 = cm[1] * A1 + cm[5] * B1 + cm[9] * C1, ly0 = cm[1] * A0 + cm[5] * B0 + cm[9] * C0;
    const float lz0 = cm[2] * A0 + cm[6] * B0 + cm[10] * C0;
    const float lu = lx1 / cm[18], lv = ly1 / cm[19];
    acc[12] = lu; acc[14] = lu * tu;
    acc[13] = lv; acc[15] = lv * tv;
    acc[16] = lz0 + lx0 * ax + ly0 * by;
#pragma unroll
    for (int t = 0; t < kCamGradTerms; ++t) {
        const float r = warp_sum(acc[t]);
        if (lane == 0) red[warp][t] = r;
    }
    __syncthreads();
    if (threadIdx.x < kCamGradTerms) {
        float r = 0.f;
        for (int w = 0; w < NW; ++w) r += red[w][threadIdx.x];
        ws[(int64_t)blockIdx.x * kCamGradTerms + threadIdx.x] = r;
    }
}

// stage 2: one warp per (camera, term); lanes stride over the block partials, fixed-order fp64 tree.
// out_stride = LF_CAMGRAD_STRIDE: the 17 terms in order (+ zero padding); out_stride = LF_CAM_STRIDE: a gradient of the
// camera block itself, terms 0..15 in place, d/d(znear) at [20], zeros elsewhere (what lf_camera_o2c_bwd consumes).
__global__ void resample_o2c_bwd_cam_finish(const float* __restrict__ ws, float* __restrict__ gcam,
                                            int blocks_per_cam, int out_stride) {
    const int n = blockIdx.x;
    const int t = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* row = gcam + (int64_t)n * out_stride;
    if (out_stride == LF_CAM_STRIDE) {
        const int c = threadIdx.x;                         // zero the columns no term lands in
        if (c < LF_CAM_STRIDE && c >= 16 && c != 20) row[c] = 0.f;
    }
    if (t >= LF_CAMGRAD_STRIDE) return;
    double a = 0.0;
    if (t < kCamGradTerms)
        for (int b = lane; b < blocks_per_cam; b += 32) a += (double)ws[((int64_t)n * blocks_per_cam + b) * kCamGradTerms + t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane != 0) return;
    if (out_stride == LF_CAM_STRIDE) {
        if (t < 16) row[t] = (float)a;
        else if (t == 16) row[20] = (float)a;
    } else {
        row[t] = (float)a;
    }
}

static int lpv_log2_for(int C, int vec) {
    int groups = (C + vec - 1) / vec, l = 0;
    while ((1 << l) < groups && l < 5) ++l;
    return l;
}

static int grid_for(int64_t total_groups, int lpv_log2) {
    int64_t threads = total_groups << lpv_log2;
    int64_t blocks = (threads + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 64;   // grid-stride beyond this
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}


template <int MODE, int LPVL, int W>
static int launch_march(const float* vol, const float* cam, float* out, int vpo, int N, int S, cudaStream_t st) {
    constexpr int LPV = 1 << LPVL, G = 32 / LPV;
    const int64_t cols = (int64_t)N * ((S + 2 * G - 1) / (2 * G)) * ((S + 3) / 4);
    // depth chunk per CTA: a multiple of LPV, as long as possible (the register-resident corners are lost at a
    // chunk start) while still giving the machine several waves of CTAs
    int KC = (S + LPV - 1) / LPV * LPV;
    while (KC > LPV && KC > 16 && cols * ((S + KC - 1) / KC) < 12ll * sm_count()) KC = ((KC / 2) + LPV - 1) / LPV * LPV;
    { const int k = option(OPT_RESAMPLE_KC); if (k >= LPV) KC = k / LPV * LPV; }
    const int64_t blocks = cols * ((S + KC - 1) / KC);
    LF_CHECK_ARG(blocks < (1ll << 31), "resample: too many columns");
    resample_march_kernel<MODE, LPVL, (W == 1 ? 3 : 2), W><<<(unsigned)blocks, 256, 0, st>>>(vol, cam, out, vpo, N, S, KC);
    LF_RETURN_LAUNCH();
}

// zero halo of a split-planar volume: per (plane, channel chunk, part) the rows yp = 0 / h+1 and the columns xp = 0 / w+1
__global__ void split_halo_zero_kernel(uint16_t* __restrict__ out, int64_t part_elems, int64_t planes_kc, int h, int w) {
    const int Wp = w + 2, HN = 2 * Wp + 2 * h;
    const int64_t PP = (int64_t)(h + 2) * Wp;
    const int64_t total = planes_kc * 2 * HN;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(idx % HN);
        const int64_t r = idx / HN;
        const int part = (int)(r & 1);
        const int64_t pk = r >> 1;
        int q;
        if (e < Wp) q = e;                                            // top row
        else if (e < 2 * Wp) q = (h + 1) * Wp + (e - Wp);             // bottom row
        else if (e < 2 * Wp + h) q = (e - 2 * Wp + 1) * Wp;           // left column
        else q = (e - 2 * Wp - h + 1) * Wp + (w + 1);                 // right column
        *reinterpret_cast<uint4*>(out + part * part_elems + (pk * PP + q) * 8) = make_uint4(0u, 0u, 0u, 0u);
    }
}

template <int LPVL>
static int launch_march_split(const float* vol, const float* cam, void* out, int vpo, int N, int S, cudaStream_t st) {
    constexpr int LPV = 1 << LPVL, G = 32 / LPV, C = 4 * LPV;
    const int64_t cols = (int64_t)N * ((S + 2 * G - 1) / (2 * G)) * ((S + 3) / 4);
    int KC = (S + LPV - 1) / LPV * LPV;
    while (KC > LPV && KC > 16 && cols * ((S + KC - 1) / KC) < 12ll * sm_count()) KC = ((KC / 2) + LPV - 1) / LPV * LPV;
    const int64_t blocks = cols * ((S + KC - 1) / KC);
    LF_CHECK_ARG(blocks < (1ll << 31), "resample: too many columns");
    const int64_t PP = (int64_t)(S + 2) * (S + 2);
    const int64_t planes_kc = (int64_t)N * S * (C / 8);
    const int64_t halo = planes_kc * 2 * (2 * (S + 2) + 2 * S);
    split_halo_zero_kernel<<<(unsigned)((halo + 255) / 256 > 4096 ? 4096 : (halo + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<uint16_t*>(out), (int64_t)N * S * C * PP, planes_kc, S, S);
    resample_march_kernel<0, LPVL, 3, 1, true><<<(unsigned)blocks, 256, 0, st>>>(vol, cam, reinterpret_cast<float*>(out), vpo, N, S, KC);
    LF_RETURN_LAUNCH();
}

// LFB200_RESAMPLE_BRICK=1 selects the brick kernel for every shape (A/B timing and the cross-check test)
static bool use_march() { return option(OPT_RESAMPLE_BRICK) != 1; }

template <int MODE>
static int launch_fwd(const float* vol, const float* cam, float* out, int vpo, int N, int C, int S, cudaStream_t st) {
    if (use_march() && (int64_t)S * S * S * (C / 4) < (1ll << 32)) {
        if (option(OPT_RESAMPLE_W) == 2) {
            if (C == 16) return launch_march<MODE, 1, 2>(vol, cam, out, vpo, N, S, st);
            if (C == 32) return launch_march<MODE, 2, 2>(vol, cam, out, vpo, N, S, st);
            if (C == 64) return launch_march<MODE, 3, 2>(vol, cam, out, vpo, N, S, st);
        }
        if (C == 16) return launch_march<MODE, 2, 1>(vol, cam, out, vpo, N, S, st);
        if (C == 32) return launch_march<MODE, 3, 1>(vol, cam, out, vpo, N, S, st);
        if (C == 64) return launch_march<MODE, 4, 1>(vol, cam, out, vpo, N, S, st);
    }
    const int64_t blocks = (int64_t)N * brick_grid(S, BX, BY, BZ).per_cam();
    LF_CHECK_ARG(blocks < (1ll << 31), "resample: too many bricks");
    if (C % 4 == 0) {
        const int l = lpv_log2_for(C, 4);
        if (C == (4 << l)) resample_fwd_kernel<MODE, 4, true><<<(unsigned)blocks, 256, 0, st>>>(vol, cam, out, vpo, N, C, S, l);
        else resample_fwd_kernel<MODE, 4, false><<<(unsigned)blocks, 256, 0, st>>>(vol, cam, out, vpo, N, C, S, l);
    } else {
        const int l = lpv_log2_for(C, 1);
        resample_fwd_kernel<MODE, 1, false><<<(unsigned)blocks, 256, 0, st>>>(vol, cam, out, vpo, N, C, S, l);
    }
    LF_RETURN_LAUNCH();
}

template <int MODE>
static int launch_bwd_vol(const float* gout, const float* cam, float* gvol, int vpo, int N, int C, int S, cudaStream_t st) {
    const int64_t total = (int64_t)N * S * S * S;
    if (C % 4 == 0) {
        const int l = lpv_log2_for(C, 4);
        resample_bwd_vol_kernel<MODE, 4><<<grid_for(total, l), 256, 0, st>>>(gout, cam, gvol, vpo, N, C, S, l);
    } else {
        const int l = lpv_log2_for(C, 1);
        resample_bwd_vol_kernel<MODE, 1><<<grid_for(total, l), 256, 0, st>>>(gout, cam, gvol, vpo, N, C, S, l);
    }
    LF_RETURN_LAUNCH();
}

static int check_common(const void* a, const void* b, const void* c, int N, int C, int S) {
    LF_CHECK_ARG(a && b && c, "resample: null pointer");
    LF_CHECK_ARG(N > 0 && C > 0 && S > 1, "resample: bad extents N=%d C=%d S=%d", N, C, S);
    LF_CHECK_ARG((int64_t)S * S * S * C < (1ll << 31), "resample: one cube must be < 2^31 elements");
    return LF_OK;
}

}  // namespace lf

using namespace lf;

extern "C" int lf_resample_o2c_fwd(const float* vol, const float* cam, float* out, int B, int N, int C, int S, void* stream) {
    if (int e = check_common(vol, cam, out, N, C, S)) return e;
    LF_CHECK_ARG(B > 0 && N % B == 0, "o2c: N=%d must be a multiple of B=%d", N, B);
    return launch_fwd<0>(vol, cam, out, N / B, N, C, S, (cudaStream_t)stream);
}

extern "C" int lf_resample_o2c_fwd_split_supported(int C, int S) {
    return (C == 16 || C == 32 || C == 64) && S > 1 && (int64_t)S * S * S * (C / 4) < (1ll << 32) &&
           (int64_t)(S + 2) * (S + 2) < (1 << 20);
}

extern "C" int lf_resample_o2c_fwd_split(const float* vol, const float* cam, void* out_split, int B, int N, int C, int S,
                                         void* stream) {
    if (int e = check_common(vol, cam, out_split, N, C, S)) return e;
    LF_CHECK_ARG(B > 0 && N % B == 0, "o2c: N=%d must be a multiple of B=%d", N, B);
    if (!lf_resample_o2c_fwd_split_supported(C, S)) {
        set_error("o2c_fwd_split: needs C in {16, 32, 64} (got C=%d S=%d)", C, S);
        return LF_EUNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 16) return launch_march_split<2>(vol, cam, out_split, N / B, N, S, st);
    if (C == 32) return launch_march_split<3>(vol, cam, out_split, N / B, N, S, st);
    return launch_march_split<4>(vol, cam, out_split, N / B, N, S, st);
}

extern "C" int lf_resample_o2c_bwd_vol(const float* gout, const float* cam, float* gvol, int B, int N, int C, int S, void* stream) {
    if (int e = check_common(gout, cam, gvol, N, C, S)) return e;
    LF_CHECK_ARG(B > 0 && N % B == 0, "o2c: N=%d must be a multiple of B=%d", N, B);
    return launch_bwd_vol<0>(gout, cam, gvol, N / B, N, C, S, (cudaStream_t)stream);
}

extern "C" int64_t lf_resample_o2c_bwd_cam_ws(int N, int S) {
    if (N <= 0 || S <= 0) return 0;
    // the larger of the brick kernel's and the marching kernel's block counts (narrowest column tile, most depth chunks)
    const int64_t brick = brick_grid(S, CBX, CBY, CBZ).per_cam();
    const int64_t march = (int64_t)((S + 3) / 4) * ((S + 1) / 2) * kBwdCamMaxChunks;
    return (int64_t)N * (brick > march ? brick : march) * kCamGradTerms;
}

// depth-marching camera gradient: C = 4 << LPVL.  LFB200_BWDCAM: 0 = this kernel (4-warp CTAs, 5 per SM), 2 = 8-warp
// CTAs at 2 per SM, 1 = the brick kernel (A/B timing and the cross-check test)
template <int LPVL>
static int launch_bwd_cam_march(const float* gout, const float* vol, const float* cam, float* gcam, float* ws,
                                int vpo, int N, int S, int out_stride, cudaStream_t st) {
    constexpr int LPV = 1 << LPVL, G = 32 / LPV;
    const int mode = option(OPT_BWDCAM);
    const int TJ = (mode == 2) ? 4 : 2;                 // CTA = 2 x TJ warps
    const int64_t cols = (int64_t)((S + 2 * G - 1) / (2 * G)) * ((S + TJ - 1) / TJ);
    // depth chunks of ~32 steps.  The split depends on S only, never on N or the machine: a camera's gradient is then
    // bit-identical whatever batch it is processed in (hypotheses sharded over GPUs reproduce the single-GPU numbers).
    int nkc = S / 32;
    nkc = nkc < 1 ? 1 : (nkc > kBwdCamMaxChunks ? kBwdCamMaxChunks : nkc);
    const int KC = ((S + nkc - 1) / nkc + LPV - 1) / LPV * LPV;
    nkc = (S + KC - 1) / KC;
    const int64_t bpc = cols * nkc;
    LF_CHECK_ARG(N * bpc < (1ll << 31), "o2c_bwd_cam: too many columns");
    const unsigned grid = (unsigned)(N * bpc);
    constexpr int D2 = LPV < 2 ? LPV : 2, D4 = LPV < 4 ? LPV : 4;
    // measured at config B (8 cameras, 64^3 x 32): 4 warps x 5 CTAs/SM (96 registers, prefetch distance 2) 133 us;
    // 8 warps x 2 CTAs/SM (128 registers, distance 4) 151 us; distance 4 at 96 registers spills (142 us)
    if (mode == 2) resample_o2c_bwd_cam_march_kernel<LPVL, 8, 2, D4><<<grid, 256, 0, st>>>(gout, vol, cam, ws, vpo, N, S, KC);
    else resample_o2c_bwd_cam_march_kernel<LPVL, 4, 5, D2><<<grid, 128, 0, st>>>(gout, vol, cam, ws, vpo, N, S, KC);
    resample_o2c_bwd_cam_finish<<<N, 32 * LF_CAMGRAD_STRIDE, 0, st>>>(ws, gcam, (int)bpc, out_stride);
    LF_RETURN_LAUNCH();
}

static int bwd_cam_impl(const float* gout, const float* vol, const float* cam, float* gcam, float* ws, int B, int N,
                        int C, int S, int out_stride, void* stream) {
    if (int e = check_common(gout, vol, cam, N, C, S)) return e;
    LF_CHECK_ARG(gcam && ws, "o2c_bwd_cam: null output/workspace");
    LF_CHECK_ARG(B > 0 && N % B == 0, "o2c: N=%d must be a multiple of B=%d", N, B);
    const int bpc = brick_grid(S, CBX, CBY, CBZ).per_cam();
    cudaStream_t st = (cudaStream_t)stream;
    if (option(OPT_BWDCAM) != 1 && (int64_t)S * S * S * (C / 4) < (1ll << 32)) {
        if (C == 16) return launch_bwd_cam_march<2>(gout, vol, cam, gcam, ws, N / B, N, S, out_stride, st);
        if (C == 32) return launch_bwd_cam_march<3>(gout, vol, cam, gcam, ws, N / B, N, S, out_stride, st);
        if (C == 64) return launch_bwd_cam_march<4>(gout, vol, cam, gcam, ws, N / B, N, S, out_stride, st);
    }
    if (C % 4 == 0) {
        const int l = lpv_log2_for(C, 4);
        resample_o2c_bwd_cam_kernel<4><<<N * bpc, 256, 0, st>>>(gout, vol, cam, ws, N / B, N, C, S, l, bpc);
    } else {
        const int l = lpv_log2_for(C, 1);
        resample_o2c_bwd_cam_kernel<1><<<N * bpc, 256, 0, st>>>(gout, vol, cam, ws, N / B, N, C, S, l, bpc);
    }
    resample_o2c_bwd_cam_finish<<<N, 32 * LF_CAMGRAD_STRIDE, 0, st>>>(ws, gcam, bpc, out_stride);
    LF_RETURN_LAUNCH();
}

extern "C" int lf_resample_o2c_bwd_cam(const float* gout, const float* vol, const float* cam, float* gcam,
                                       float* ws, int B, int N, int C, int S, void* stream) {
    return bwd_cam_impl(gout, vol, cam, gcam, ws, B, N, C, S, LF_CAMGRAD_STRIDE, stream);
}

extern "C" int lf_resample_o2c_bwd_cam_block(const float* gout, const float* vol, const float* cam, float* gblock,
                                             float* ws, int B, int N, int C, int S, void* stream) {
    return bwd_cam_impl(gout, vol, cam, gblock, ws, B, N, C, S, LF_CAM_STRIDE, stream);
}

extern "C" int lf_resample_c2o_fwd(const float* vol, const float* cam, float* out, int V, int C, int S, void* stream) {
    if (int e = check_common(vol, cam, out, V, C, S)) return e;
    return launch_fwd<1>(vol, cam, out, 1, V, C, S, (cudaStream_t)stream);
}

extern "C" int lf_resample_c2o_bwd_vol(const float* gout, const float* cam, float* gvol, int V, int C, int S, void* stream) {
    if (int e = check_common(gout, cam, gvol, V, C, S)) return e;
    return launch_bwd_vol<1>(gout, cam, gvol, 1, V, C, S, (cudaStream_t)stream);
}
