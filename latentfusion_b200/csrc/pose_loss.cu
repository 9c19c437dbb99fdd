// Fused pose-loss head.
//
// Replaces, per refinement iteration, the reference's
//   interpret_logits (recon/models.py:455-484: tanh, sigmoid, mask gate)
//   Camera.denormalize_depth (modules/geometry.py:555-558)
//   Camera.uncrop x2 (geometry.py:261-285: 2-D grid_sample to the 640x480 frame, nearest for depth,
//                     bilinear for mask logits, padding_mode='border')
//   default_pose_loss (pose/estimation.py:70-118) + pose/utils.py:81-117 reductions
// i.e. ~100 elementwise/reduction launches over [N,1,480,640] intermediates plus two
// grid_sampler_2d_backward launches that serialise on border-pixel atomics (1.5 ms each on B200),
// by two passes over the full frame that never materialise a full-frame tensor:
//   pass 1: per-hypothesis partial sums  ->  the four loss terms
//   pass 2: d(terms)/d(depth logits, mask logits, viewport, translation_z)
#include "common.cuh"

namespace lf {

constexpr int kSums = 6;   // A=sum dl1, B=sum dl1*pm*tm, C=sum pm*tm, D=sum pm, E=sum pm*tm*valid, F=sum bce

struct LossGeom {
    int n, p, width, height;
    float range, base_off;    // z = (tanh(dl)+1)/2 * gate * range + (tz + base_off)
    int ps, hs, tzs;          // strides in floats: between crop pixels of a logit map, between hypotheses, between tz entries
    int premask;              // coarse search (PoseEstimator._render_observation, estimation.py:187-197): the crop's metric
                              // depth is multiplied by the crop's own sigmoid(mask) before it is pasted into the frame
};

struct PixelSample {
    // nearest tap (depth) and bilinear taps (mask logits) of one full-frame pixel in the P x P crop
    int near_idx;
    int i00, i01, i10, i11;
    float w00, w01, w10, w11;
    float fx, fy;             // bilinear fractions
    float mx, my;             // d(ix)/d(unclipped ix): 1 inside, 0 where border-clamped
    int x0in, y0in;           // whether the +1 taps are in range (weight is 0 otherwise)
};

// ATen grid_sampler (align_corners=False, padding border): unnormalize, clip to [0, P-1]
__device__ __forceinline__ float clip_coord(float g, int P, float& mult) {
    float ix = ((g + 1.f) * (float)P - 1.f) / 2.f;
    const float mxv = (float)(P - 1);
    if (ix <= 0.f) { ix = 0.f; mult = 0.f; }
    else if (ix >= mxv) { ix = mxv; mult = 0.f; }
    else mult = 1.f;
    return ix;
}

__device__ __forceinline__ PixelSample make_sample(float X, float Y, const float* vp, int P, int ps) {
    PixelSample s;
    const float vw = vp[2] - vp[0], vh = vp[3] - vp[1];
    const float gx = (X - vp[0]) / vw * 2.f - 1.f;       // geometry.py:281-282
    const float gy = (Y - vp[1]) / vh * 2.f - 1.f;
    float ix = clip_coord(gx, P, s.mx), iy = clip_coord(gy, P, s.my);
    // degenerate optimised viewport (zero width / NaN): the reference only propagates NaNs; keep the indices in range
    if (!isfinite(ix)) { ix = 0.f; s.mx = 0.f; }
    if (!isfinite(iy)) { iy = 0.f; s.my = 0.f; }
    const int xn = (int)nearbyintf(ix), yn = (int)nearbyintf(iy);
    s.near_idx = (yn * P + xn) * ps;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    s.fx = ix - fx0; s.fy = iy - fy0;
    s.x0in = (x0 + 1 < P); s.y0in = (y0 + 1 < P);
    const int x1 = s.x0in ? x0 + 1 : x0, y1 = s.y0in ? y0 + 1 : y0;
    s.i00 = (y0 * P + x0) * ps; s.i01 = (y0 * P + x1) * ps; s.i10 = (y1 * P + x0) * ps; s.i11 = (y1 * P + x1) * ps;
    const float wx1 = s.x0in ? s.fx : 0.f, wy1 = s.y0in ? s.fy : 0.f;
    const float wx0 = 1.f - s.fx, wy0 = 1.f - s.fy;
    s.w00 = wx0 * wy0; s.w01 = wx1 * wy0; s.w10 = wx0 * wy1; s.w11 = wx1 * wy1;
    return s;
}

__device__ __forceinline__ float sigmoid_(float x) { return 1.f / (1.f + expf(-x)); }

struct PixelTerms { float z, pm, pd, dl1, td, tm, valid, ml, gate, th; };

__device__ __forceinline__ PixelTerms eval_pixel(const PixelSample& s, const float* __restrict__ dl,
                                                 const float* __restrict__ ml, float tdepth, float tmask,
                                                 float range, float base, int premask = 0) {
    PixelTerms t;
    t.th = tanhf(dl[s.near_idx]);
    t.gate = sigmoid_(ml[s.near_idx]) > 0.5f ? 1.f : 0.f;          // apply_mask (models.py:478-481)
    t.z = (t.th + 1.f) * 0.5f * t.gate * range + base;
    if (premask) t.z *= sigmoid_(ml[s.near_idx]);
    t.ml = s.w00 * ml[s.i00] + s.w01 * ml[s.i01] + s.w10 * ml[s.i10] + s.w11 * ml[s.i11];
    t.pm = sigmoid_(t.ml);
    t.pd = t.z * t.pm;
    t.valid = ((tdepth == 0.f) && (tmask > 0.1f)) ? 0.f : 1.f;
    t.tm = tmask;
    t.td = tdepth * tmask;                                          // Observation.prepare()
    t.dl1 = fabsf(t.pd - t.td) * t.valid;
    return t;
}

__global__ void __launch_bounds__(256)
pose_loss_sums_kernel(const LossGeom g, const float* __restrict__ dlog, const float* __restrict__ mlog,
                      const float* __restrict__ vp, const float* __restrict__ tz,
                      const float* __restrict__ tdepth, const float* __restrict__ tmask, float* __restrict__ sums) {
    const int n = blockIdx.y;
    const int HW = g.width * g.height;
    const float* dl = dlog + (size_t)n * g.hs;
    const float* ml = mlog + (size_t)n * g.hs;
    const float base = tz[n * g.tzs] + g.base_off;
    float acc[kSums] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int px = blockIdx.x * blockDim.x + threadIdx.x; px < HW; px += gridDim.x * blockDim.x) {
        const int Y = px / g.width, X = px - Y * g.width;
        const PixelSample s = make_sample((float)X, (float)Y, vp + 4 * n, g.p, g.ps);
        const PixelTerms t = eval_pixel(s, dl, ml, tdepth[px], tmask[px], g.range, base, g.premask);
        acc[0] += t.dl1;
        acc[1] += t.dl1 * t.pm * t.tm;
        acc[2] += t.pm * t.tm;
        acc[3] += t.pm;
        acc[4] += t.pm * t.tm * t.valid;
        acc[5] += fmaxf(t.ml, 0.f) - t.ml * t.tm + log1pf(expf(-fabsf(t.ml)));   // BCE-with-logits
    }
    __shared__ float red[8][kSums];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kSums; ++k) {
        const float r = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = r;
    }
    __syncthreads();
    if (threadIdx.x < kSums) {
        float r = 0.f;
        for (int w = 0; w < 8; ++w) r += red[w][threadIdx.x];
        atomicAdd(sums + n * 8 + threadIdx.x, r);
    }
}

// terms[n] = (ov_depth, depth, iou, mask); sums[n][6] = sum target_mask*valid (set by the first kernel below)
__global__ void pose_loss_terms_kernel(const LossGeom g, const float* __restrict__ sums, float* __restrict__ terms) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= g.n) return;
    const float* s = sums + n * 8;
    const float HW = (float)(g.width * g.height);
    const float G = s[6];
    const float U = s[3] + G - s[4];
    terms[n * 4 + 0] = fmaxf(s[1], 1e-5f) / fmaxf(s[2], 1e-4f);            // pose/utils.py:111-117
    terms[n * 4 + 1] = s[0] / HW;
    terms[n * 4 + 2] = logf(fmaxf(U, 1e-4f)) - logf(fmaxf(s[4], 1e-4f));  // pose/utils.py:99-108
    terms[n * 4 + 3] = s[5] / HW;
}

__global__ void __launch_bounds__(256)
target_sum_kernel(const LossGeom g, const float* __restrict__ tdepth, const float* __restrict__ tmask, float* __restrict__ sums) {
    const int HW = g.width * g.height;
    float a = 0.f;
    for (int px = blockIdx.x * blockDim.x + threadIdx.x; px < HW; px += gridDim.x * blockDim.x) {
        const float tm = tmask[px];
        a += ((tdepth[px] == 0.f) && (tm > 0.1f)) ? 0.f : tm;
    }
    a = warp_sum(a);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float r = 0.f;
        for (int w = 0; w < 8; ++w) r += red[w];
        for (int n = 0; n < g.n; ++n) atomicAdd(sums + n * 8 + 6, r);
    }
}

// one atomic per warp when every lane hits the same address (pixels clamped to the crop border), else per lane
__device__ __forceinline__ void warp_atomic_add(float* base, int idx, float v) {
    const int idx0 = __shfl_sync(0xffffffffu, idx, 0);
    if (__all_sync(0xffffffffu, idx == idx0)) {
        v = warp_sum(v);
        if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(base + idx0, v);
    } else if (v != 0.f) {
        atomicAdd(base + idx, v);
    }
}

__global__ void __launch_bounds__(256, 4)
pose_loss_bwd_kernel(const LossGeom g, const float* __restrict__ dlog, const float* __restrict__ mlog,
                     const float* __restrict__ vp, const float* __restrict__ tz,
                     const float* __restrict__ tdepth, const float* __restrict__ tmask,
                     const float* __restrict__ sums, const float* __restrict__ gterms,
                     float* __restrict__ g_dl, float* __restrict__ g_ml, float* __restrict__ g_vp, float* __restrict__ g_tz) {
    const int n = blockIdx.y;
    const int HW = g.width * g.height;
    const float* dl = dlog + (size_t)n * g.hs;
    const float* ml = mlog + (size_t)n * g.hs;
    float* gdl = g_dl + (size_t)n * g.hs;
    float* gml = g_ml + (size_t)n * g.hs;
    const float base = tz[n * g.tzs] + g.base_off;
    const float* s = sums + n * 8;
    const float fHW = (float)HW;
    // d(total)/d(sums) from d(total)/d(terms)
    const float g_ov = gterms[n * 4 + 0], g_dep = gterms[n * 4 + 1], g_iou = gterms[n * 4 + 2], g_msk = gterms[n * 4 + 3];
    const float Cc = fmaxf(s[2], 1e-4f), Bc = fmaxf(s[1], 1e-5f);
    const float U = s[3] + s[6] - s[4];
    const float dA = g_dep / fHW;
    const float dB = (s[1] > 1e-5f) ? g_ov / Cc : 0.f;
    const float dC = (s[2] > 1e-4f) ? -g_ov * Bc / (Cc * Cc) : 0.f;
    const float dU = (U > 1e-4f) ? g_iou / U : 0.f;
    const float dD = dU;
    const float dE = ((s[4] > 1e-4f) ? -g_iou / s[4] : 0.f) - dU;
    const float dF = g_msk / fHW;
    const float vx0 = vp[4 * n], vy0 = vp[4 * n + 1], vw = vp[4 * n + 2] - vx0, vh = vp[4 * n + 3] - vy0;
    const float inv_vw = 1.f / vw, inv_vh = 1.f / vh, two_inv_vw = 2.f * inv_vw, two_inv_vh = 2.f * inv_vh;
    const float half_p = (float)g.p * 0.5f;            // d ix / d gx

    float a_tz = 0.f, a_vp[4] = {0.f, 0.f, 0.f, 0.f};
    // full warps only: the aggregation helper uses warp-wide votes
    const int HW_pad = (HW + 31) & ~31;
    for (int px = blockIdx.x * blockDim.x + threadIdx.x; px < HW_pad; px += gridDim.x * blockDim.x) {
        const bool live = px < HW;
        const int pc = live ? px : HW - 1;
        const int Y = pc / g.width, X = pc - Y * g.width;
        const PixelSample smp = make_sample((float)X, (float)Y, vp + 4 * n, g.p, g.ps);
        const PixelTerms t = eval_pixel(smp, dl, ml, tdepth[pc], tmask[pc], g.range, base);
        float d_dl1 = dA + dB * t.pm * t.tm;
        float d_pm = dB * t.dl1 * t.tm + dC * t.tm + dD + dE * t.tm * t.valid;
        const float diff = t.pd - t.td;
        const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        const float d_pd = d_dl1 * t.valid * sgn;
        const float d_z = d_pd * t.pm;
        d_pm += d_pd * t.z;
        float d_ml = d_pm * t.pm * (1.f - t.pm) + dF * (t.pm - t.tm);
        float d_dlog = d_z * (1.f - t.th * t.th) * 0.5f * t.gate * g.range;
        if (!live) { d_ml = 0.f; d_dlog = 0.f; }
        a_tz += live ? d_z : 0.f;
        // bilinear taps of the mask logits
        warp_atomic_add(gml, smp.i00, d_ml * smp.w00);
        warp_atomic_add(gml, smp.i01, d_ml * smp.w01);
        warp_atomic_add(gml, smp.i10, d_ml * smp.w10);
        warp_atomic_add(gml, smp.i11, d_ml * smp.w11);
        warp_atomic_add(gdl, smp.near_idx, d_dlog);
        // d(ml_full)/d(ix, iy) -> viewport (ATen grid_sampler_2d_backward: gix uses the in-range taps only)
        const float m00 = ml[smp.i00], m01 = smp.x0in ? ml[smp.i01] : 0.f;
        const float m10 = smp.y0in ? ml[smp.i10] : 0.f, m11 = (smp.x0in && smp.y0in) ? ml[smp.i11] : 0.f;
        const float wy0 = 1.f - smp.fy, wy1 = smp.y0in ? smp.fy : 0.f;
        const float wx0 = 1.f - smp.fx, wx1 = smp.x0in ? smp.fx : 0.f;
        const float dml_dix = (smp.x0in ? (m01 - m00) : -m00) * wy0 + ((smp.x0in ? m11 : 0.f) - m10) * wy1;
        const float dml_diy = (smp.y0in ? (m10 - m00) : -m00) * wx0 + ((smp.y0in ? m11 : 0.f) - m01) * wx1;
        const float gix = d_ml * dml_dix * smp.mx * half_p;
        const float giy = d_ml * dml_diy * smp.my * half_p;
        // gx = (X - vx0)/vw*2 - 1:  d gx / d vx0 = 2*(rx - 1)/vw,  d gx / d vx1 = -2*rx/vw  with rx = (X - vx0)/vw
        const float rx = ((float)X - vx0) * inv_vw, ry = ((float)Y - vy0) * inv_vh;
        const float tx = gix * two_inv_vw, ty = giy * two_inv_vh;
        a_vp[0] += tx * (rx - 1.f);
        a_vp[2] -= tx * rx;
        a_vp[1] += ty * (ry - 1.f);
        a_vp[3] -= ty * ry;
    }
    __shared__ float red[8][5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float v5[5] = {a_vp[0], a_vp[1], a_vp[2], a_vp[3], a_tz};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const float r = warp_sum(v5[k]);
        if (lane == 0) red[warp][k] = r;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        float r = 0.f;
        for (int w = 0; w < 8; ++w) r += red[w][threadIdx.x];
        if (threadIdx.x < 4) atomicAdd(g_vp + 4 * n + threadIdx.x, r);
        else atomicAdd(g_tz + n * g.tzs, r);
    }
}

// blocks along x for a (bx, n) grid of grid-stride CTAs: as many as are co-resident (occupancy x SMs), so that the
// launch is ONE balanced wave (600 CTAs on 444 slots ran as two waves, the second one nearly empty)
template <typename K>
static int blocks_x(K kernel, int n, int HW) {
    static int per_sm = 0;                      // per kernel instantiation (K is a distinct function type per kernel)
    if (per_sm == 0) {
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, 256, 0) != cudaSuccess || v < 1) v = 2;
        per_sm = v;
    }
    int bx = per_sm * sm_count() / max(1, n);
    bx = max(1, min(bx, (HW + 255) / 256));
    return bx;
}

static int loss_geom(const lf_loss_desc* d, LossGeom& g) {
    LF_CHECK_ARG(d != nullptr, "pose_loss: null descriptor");
    LF_CHECK_ARG(d->n > 0 && d->p > 1 && d->width > 0 && d->height > 0, "pose_loss: bad extents");
    g.n = d->n; g.p = d->p; g.width = d->width; g.height = d->height;
    g.range = 2.f * d->z_span + 2.f * d->eps;          // (zfar + eps) - (znear - eps)
    g.base_off = -d->z_span - d->eps;                  // znear - eps = tz - z_span - eps
    g.premask = 0;
    LF_CHECK_ARG(d->pix_stride >= 0 && d->hyp_stride >= 0 && d->tz_stride >= 0, "pose_loss: negative stride");
    g.ps = d->pix_stride > 0 ? d->pix_stride : 1;
    g.hs = d->hyp_stride > 0 ? d->hyp_stride : d->p * d->p * g.ps;
    g.tzs = d->tz_stride > 0 ? d->tz_stride : 1;
    LF_CHECK_ARG((int64_t)g.hs >= (int64_t)(d->p * d->p - 1) * g.ps + 1, "pose_loss: hypothesis stride smaller than one logit map");
    return LF_OK;
}

}  // namespace lf

using namespace lf;

extern "C" int lf_pose_loss_fwd(const lf_loss_desc* desc, const float* depth_logits, const float* mask_logits,
                                const float* viewport, const float* tz, const float* target_depth,
                                const float* target_mask, float* sums, float* terms, void* stream) {
    LossGeom g;
    if (int e = loss_geom(desc, g)) return e;
    LF_CHECK_ARG(depth_logits && mask_logits && viewport && tz && target_depth && target_mask && sums && terms,
                 "pose_loss_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(sums, 0, sizeof(float) * 8 * g.n, st);
    const int HW = g.width * g.height;
    const int bx = blocks_x(pose_loss_sums_kernel, g.n, HW);
    target_sum_kernel<<<min(2 * sm_count(), (HW + 255) / 256), 256, 0, st>>>(g, target_depth, target_mask, sums);
    pose_loss_sums_kernel<<<dim3(bx, g.n), 256, 0, st>>>(g, depth_logits, mask_logits, viewport, tz, target_depth, target_mask, sums);
    pose_loss_terms_kernel<<<(g.n + 63) / 64, 64, 0, st>>>(g, sums, terms);
    LF_RETURN_LAUNCH();
}

// Forward-only scoring for the coarse pose search (CrossEntropyPoseEstimator, reference estimation.py:187-197 +
// :70-118): same four terms, with the search's extra crop-space mask factor on the rendered depth.
extern "C" int lf_pose_loss_search_fwd(const lf_loss_desc* desc, const float* depth_logits, const float* mask_logits,
                                       const float* viewport, const float* tz, const float* target_depth,
                                       const float* target_mask, float* sums, float* terms, void* stream) {
    LossGeom g;
    if (int e = loss_geom(desc, g)) return e;
    g.premask = 1;
    LF_CHECK_ARG(depth_logits && mask_logits && viewport && tz && target_depth && target_mask && sums && terms,
                 "pose_loss_search_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(sums, 0, sizeof(float) * 8 * g.n, st);
    const int HW = g.width * g.height;
    const int bx = blocks_x(pose_loss_sums_kernel, g.n, HW);
    target_sum_kernel<<<min(2 * sm_count(), (HW + 255) / 256), 256, 0, st>>>(g, target_depth, target_mask, sums);
    pose_loss_sums_kernel<<<dim3(bx, g.n), 256, 0, st>>>(g, depth_logits, mask_logits, viewport, tz, target_depth, target_mask, sums);
    pose_loss_terms_kernel<<<(g.n + 63) / 64, 64, 0, st>>>(g, sums, terms);
    LF_RETURN_LAUNCH();
}

extern "C" int lf_pose_loss_bwd(const lf_loss_desc* desc, const float* depth_logits, const float* mask_logits,
                                const float* viewport, const float* tz, const float* target_depth,
                                const float* target_mask, const float* sums, const float* grad_terms,
                                float* grad_depth_logits, float* grad_mask_logits, float* grad_viewport,
                                float* grad_tz, void* stream) {
    LossGeom g;
    if (int e = loss_geom(desc, g)) return e;
    LF_CHECK_ARG(depth_logits && mask_logits && viewport && tz && target_depth && target_mask && sums && grad_terms &&
                 grad_depth_logits && grad_mask_logits && grad_viewport && grad_tz, "pose_loss_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    // dense layout: the gradients are zero-filled here; strided layout (maps interleaved in one tensor, tz a column of
    // the translation): the caller passes zero-filled tensors of the inputs' layout
    if (g.ps == 1 && g.hs == g.p * g.p) {
        const size_t crop = sizeof(float) * (size_t)g.n * g.p * g.p;
        cudaMemsetAsync(grad_depth_logits, 0, crop, st);
        cudaMemsetAsync(grad_mask_logits, 0, crop, st);
    }
    cudaMemsetAsync(grad_viewport, 0, sizeof(float) * 4 * g.n, st);
    if (g.tzs == 1) cudaMemsetAsync(grad_tz, 0, sizeof(float) * g.n, st);
    const int HW = g.width * g.height;
    const int bx = blocks_x(pose_loss_bwd_kernel, g.n, HW);
    pose_loss_bwd_kernel<<<dim3(bx, g.n), 256, 0, st>>>(g, depth_logits, mask_logits, viewport, tz, target_depth,
                                                      target_mask, sums, grad_terms, grad_depth_logits,
                                                      grad_mask_logits, grad_viewport, grad_tz);
    LF_RETURN_LAUNCH();
}
