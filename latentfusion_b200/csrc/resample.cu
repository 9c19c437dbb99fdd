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
        acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
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
    int64_t o[8];   // element offsets (before channel) in ATen order: tnw tne tsw tse bnw bne bsw bse
    float w[8];
};

__device__ __forceinline__ CornerOfs corner_offsets(const Samp& s, int S, int C) {
    CornerOfs c;
    const int x1 = min(s.x0 + 1, S - 1), y1 = min(s.y0 + 1, S - 1), z1 = min(s.z0 + 1, S - 1);
    // a +1 corner that falls outside only happens at ix == S-1 where its weight is exactly 0
    const int64_t zs0 = (int64_t)s.z0 * S, zs1 = (int64_t)z1 * S;
    c.o[0] = ((zs0 + s.y0) * S + s.x0) * C;  c.w[0] = s.wx0 * s.wy0 * s.wz0;
    c.o[1] = ((zs0 + s.y0) * S + x1) * C;    c.w[1] = s.wx1 * s.wy0 * s.wz0;
    c.o[2] = ((zs0 + y1) * S + s.x0) * C;    c.w[2] = s.wx0 * s.wy1 * s.wz0;
    c.o[3] = ((zs0 + y1) * S + x1) * C;      c.w[3] = s.wx1 * s.wy1 * s.wz0;
    c.o[4] = ((zs1 + s.y0) * S + s.x0) * C;  c.w[4] = s.wx0 * s.wy0 * s.wz1;
    c.o[5] = ((zs1 + s.y0) * S + x1) * C;    c.w[5] = s.wx1 * s.wy0 * s.wz1;
    c.o[6] = ((zs1 + y1) * S + s.x0) * C;    c.w[6] = s.wx0 * s.wy1 * s.wz1;
    c.o[7] = ((zs1 + y1) * S + x1) * C;      c.w[7] = s.wx1 * s.wy1 * s.wz1;
    return c;
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

template <int MODE, int VEC>
__global__ void __launch_bounds__(256)
resample_fwd_kernel(const float* __restrict__ vol, const float* __restrict__ cam, float* __restrict__ out,
                    int views_per_obj, int N, int C, int S, int lpv_log2) {
    const int64_t S3 = (int64_t)S * S * S;
    const BrickGrid bg = brick_grid(S, BX, BY, BZ);
    const int n = blockIdx.x / bg.per_cam();
    int b = blockIdx.x - n * bg.per_cam();
    const int bi = b % bg.nbx; b /= bg.nbx;
    const int bj = b % bg.nby; const int bk = b / bg.nby;
    const int lpv = 1 << lpv_log2;
    const int sub = threadIdx.x & (lpv - 1);
    const int grp = threadIdx.x >> lpv_log2;
    const int ngrp = blockDim.x >> lpv_log2;

    __shared__ float cm[LF_CAM_STRIDE];
    if (threadIdx.x < LF_CAM_STRIDE) cm[threadIdx.x] = cam[(int64_t)n * LF_CAM_STRIDE + threadIdx.x];
    __syncthreads();

    const float* vb = vol + (int64_t)(MODE == 0 ? n / views_per_obj : n) * S3 * C;
    for (int lv = grp; lv < BX * BY * BZ; lv += ngrp) {
        const int i = bi * BX + (lv % BX);
        const int j = bj * BY + (lv / BX) % BY;
        const int k = bk * BZ + lv / (BX * BY);
        if (i >= S || j >= S || k >= S) continue;
        float gx, gy, gz;
        gen_grid<MODE>(cm, S, i, j, k, gx, gy, gz);
        const Samp s = make_samp(gx, gy, gz, S);
        const CornerOfs co = corner_offsets(s, S, C);
        float* ob = out + ((int64_t)n * S3 + ((int64_t)k * S + j) * S + i) * C;
        for (int c = sub * VEC; c < C; c += lpv * VEC) {
            typename Vec<VEC>::T val[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) val[q] = Vec<VEC>::load(vb + co.o[q] + c);
            typename Vec<VEC>::T acc = Vec<VEC>::zero();
#pragma unroll
            for (int q = 0; q < 8; ++q) Vec<VEC>::fma(acc, co.w[q], val[q]);
            Vec<VEC>::store(ob + c, acc);
        }
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
constexpr int CBX = 16, CBY = 16, CBZ = 8;  // bwd_cam brick: 2048 voxels per block
constexpr int kBwdCamChunk = CBX * CBY * CBZ;

template <int VEC>
__global__ void __launch_bounds__(256)
resample_o2c_bwd_cam_kernel(const float* __restrict__ gout, const float* __restrict__ vol,
                            const float* __restrict__ cam, float* __restrict__ ws,
                            int views_per_obj, int N, int C, int S, int lpv_log2, int blocks_per_cam) {
    const int S3 = S * S * S;
    const int n = blockIdx.x / blocks_per_cam;
    const int b = blockIdx.x - n * blocks_per_cam;
    const int lpv = 1 << lpv_log2;
    const int sub = threadIdx.x & (lpv - 1);
    const int grp = threadIdx.x >> lpv_log2;
    const int ngrp = blockDim.x >> lpv_log2;

    __shared__ float cm[LF_CAM_STRIDE];
    if (threadIdx.x < LF_CAM_STRIDE) cm[threadIdx.x] = cam[(int64_t)n * LF_CAM_STRIDE + threadIdx.x];
    __syncthreads();

    float acc[kCamGradTerms];
#pragma unroll
    for (int t = 0; t < kCamGradTerms; ++t) acc[t] = 0.f;

    const float* vb = vol + (int64_t)(n / views_per_obj) * S3 * C;
    const BrickGrid bg = brick_grid(S, CBX, CBY, CBZ);
    int bb_ = b;
    const int bi = bb_ % bg.nbx; bb_ /= bg.nbx;
    const int bj = bb_ % bg.nby; const int bk = bb_ / bg.nby;
    // warp-uniform trip count (the shuffles below use the full mask); voxels outside the cube are
    // clamped to a valid one and their result dropped
    for (int lv0 = 0; lv0 < kBwdCamChunk; lv0 += ngrp) {
        const int lv = lv0 + grp;
        int i = bi * CBX + (lv % CBX), j = bj * CBY + (lv / CBX) % CBY, k = bk * CBZ + lv / (CBX * CBY);
        const bool valid = (lv < kBwdCamChunk) && i < S && j < S && k < S;
        i = min(i, S - 1); j = min(j, S - 1); k = min(k, S - 1);
        const int p = (k * S + j) * S + i;
        // --- forward recompute of the grid (same op order as gen_grid<0>) keeping intermediates
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
        const CornerOfs co = corner_offsets(s, S, C);
        const float* gb = gout + ((int64_t)n * S3 + p) * C;

        // d(sample)/d(ix,iy,iz) contracted with grad_out over this lane's channels
        float dx = 0.f, dy = 0.f, dz = 0.f;
        for (int c = sub * VEC; c < C; c += lpv * VEC) {
            typename Vec<VEC>::T val[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) val[q] = Vec<VEC>::load(vb + co.o[q] + c);
            const typename Vec<VEC>::T g = Vec<VEC>::load(gb + c);
            // corner order: q = zbit*4 + ybit*2 + xbit
            const float e01 = Vec<VEC>::dot(g, Vec<VEC>::sub(val[1], val[0]));
            const float e23 = Vec<VEC>::dot(g, Vec<VEC>::sub(val[3], val[2]));
            const float e45 = Vec<VEC>::dot(g, Vec<VEC>::sub(val[5], val[4]));
            const float e67 = Vec<VEC>::dot(g, Vec<VEC>::sub(val[7], val[6]));
            dx += e01 * s.wy0 * s.wz0 + e23 * s.wy1 * s.wz0 + e45 * s.wy0 * s.wz1 + e67 * s.wy1 * s.wz1;
            const float f02 = Vec<VEC>::dot(g, Vec<VEC>::sub(val[2], val[0]));
            const float f13 = Vec<VEC>::dot(g, Vec<VEC>::sub(val[3], val[1]));
            const float f46 = Vec<VEC>::dot(g, Vec<VEC>::sub(val[6], val[4]));
            const float f57 = Vec<VEC>::dot(g, Vec<VEC>::sub(val[7], val[5]));
            dy += f02 * s.wx0 * s.wz0 + f13 * s.wx1 * s.wz0 + f46 * s.wx0 * s.wz1 + f57 * s.wx1 * s.wz1;
            const float h04 = Vec<VEC>::dot(g, Vec<VEC>::sub(val[4], val[0]));
            const float h15 = Vec<VEC>::dot(g, Vec<VEC>::sub(val[5], val[1]));
            const float h26 = Vec<VEC>::dot(g, Vec<VEC>::sub(val[6], val[2]));
            const float h37 = Vec<VEC>::dot(g, Vec<VEC>::sub(val[7], val[3]));
            dz += h04 * s.wx0 * s.wy0 + h15 * s.wx1 * s.wy0 + h26 * s.wx0 * s.wy1 + h37 * s.wx1 * s.wy1;
        }
        for (int o = lpv >> 1; o > 0; o >>= 1) {
            dx += __shfl_xor_sync(0xffffffffu, dx, o);
            dy += __shfl_xor_sync(0xffffffffu, dy, o);
            dz += __shfl_xor_sync(0xffffffffu, dz, o);
        }
        if (sub == 0 && valid) {
            // chain: grid -> object coords -> (M, x, y, z) -> (viewport, znear)
            const float gox = dx * s.mx / half, goy = dy * s.my / half, goz = dz * s.mz / half;
            acc[0] += gox * x; acc[1] += gox * y; acc[2] += gox * z; acc[3] += gox;
            acc[4] += goy * x; acc[5] += goy * y; acc[6] += goy * z; acc[7] += goy;
            acc[8] += goz * x; acc[9] += goz * y; acc[10] += goz * z; acc[11] += goz;
            const float lx = cm[0] * gox + cm[4] * goy + cm[8] * goz;
            const float ly = cm[1] * gox + cm[5] * goy + cm[9] * goz;
            const float lz = cm[2] * gox + cm[6] * goy + cm[10] * goz;
            const float lu = lx * z / cm[18];     // dL/du
            const float lv = ly * z / cm[19];     // dL/dv
            acc[12] += lu; acc[14] += lu * tu;
            acc[13] += lv; acc[15] += lv * tv;
            acc[16] += lz + lx * a + ly * bb;     // dL/dznear  (z = tz*z_span + znear)
        }
    }

    // block reduction
    __shared__ float red[8][kCamGradTerms];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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

__global__ void resample_o2c_bwd_cam_finish(const float* __restrict__ ws, float* __restrict__ gcam,
                                            int blocks_per_cam) {
    const int n = blockIdx.x;
    const int t = threadIdx.x;
    if (t >= LF_CAMGRAD_STRIDE) return;
    float r = 0.f;
    if (t < kCamGradTerms) {
        double a = 0.0;
        for (int b = 0; b < blocks_per_cam; ++b) a += (double)ws[((int64_t)n * blocks_per_cam + b) * kCamGradTerms + t];
        r = (float)a;
    }
    gcam[(int64_t)n * LF_CAMGRAD_STRIDE + t] = r;
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

template <int MODE>
static int launch_fwd(const float* vol, const float* cam, float* out, int vpo, int N, int C, int S, cudaStream_t st) {
    const int64_t blocks = (int64_t)N * brick_grid(S, BX, BY, BZ).per_cam();
    LF_CHECK_ARG(blocks < (1ll << 31), "resample: too many bricks");
    if (C % 4 == 0) {
        const int l = lpv_log2_for(C, 4);
        resample_fwd_kernel<MODE, 4><<<(unsigned)blocks, 256, 0, st>>>(vol, cam, out, vpo, N, C, S, l);
    } else {
        const int l = lpv_log2_for(C, 1);
        resample_fwd_kernel<MODE, 1><<<(unsigned)blocks, 256, 0, st>>>(vol, cam, out, vpo, N, C, S, l);
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

extern "C" int lf_resample_o2c_bwd_vol(const float* gout, const float* cam, float* gvol, int B, int N, int C, int S, void* stream) {
    if (int e = check_common(gout, cam, gvol, N, C, S)) return e;
    LF_CHECK_ARG(B > 0 && N % B == 0, "o2c: N=%d must be a multiple of B=%d", N, B);
    return launch_bwd_vol<0>(gout, cam, gvol, N / B, N, C, S, (cudaStream_t)stream);
}

extern "C" int64_t lf_resample_o2c_bwd_cam_ws(int N, int S) {
    if (N <= 0 || S <= 0) return 0;
    return (int64_t)N * brick_grid(S, CBX, CBY, CBZ).per_cam() * kCamGradTerms;
}

extern "C" int lf_resample_o2c_bwd_cam(const float* gout, const float* vol, const float* cam, float* gcam,
                                       float* ws, int B, int N, int C, int S, void* stream) {
    if (int e = check_common(gout, vol, cam, N, C, S)) return e;
    LF_CHECK_ARG(gcam && ws, "o2c_bwd_cam: null output/workspace");
    LF_CHECK_ARG(B > 0 && N % B == 0, "o2c: N=%d must be a multiple of B=%d", N, B);
    const int bpc = brick_grid(S, CBX, CBY, CBZ).per_cam();
    cudaStream_t st = (cudaStream_t)stream;
    if (C % 4 == 0) {
        const int l = lpv_log2_for(C, 4);
        resample_o2c_bwd_cam_kernel<4><<<N * bpc, 256, 0, st>>>(gout, vol, cam, ws, N / B, N, C, S, l, bpc);
    } else {
        const int l = lpv_log2_for(C, 1);
        resample_o2c_bwd_cam_kernel<1><<<N * bpc, 256, 0, st>>>(gout, vol, cam, ws, N / B, N, C, S, l, bpc);
    }
    resample_o2c_bwd_cam_finish<<<N, 32, 0, st>>>(ws, gcam, bpc);
    LF_RETURN_LAUNCH();
}

extern "C" int lf_resample_c2o_fwd(const float* vol, const float* cam, float* out, int V, int C, int S, void* stream) {
    if (int e = check_common(vol, cam, out, V, C, S)) return e;
    return launch_fwd<1>(vol, cam, out, 1, V, C, S, (cudaStream_t)stream);
}

extern "C" int lf_resample_c2o_bwd_vol(const float* gout, const float* cam, float* gvol, int V, int C, int S, void* stream) {
    if (int e = check_common(gout, cam, gvol, V, C, S)) return e;
    return launch_bwd_vol<1>(gout, cam, gvol, 1, V, C, S, (cudaStream_t)stream);
}
