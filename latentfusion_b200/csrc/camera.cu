// Camera algebra on device: the ~200 tiny ATen launches per iteration that turn the
// ten learnable floats of every hypothesis into the resampler's constant block — qexp -> normalize x2 ->
// quat_to_mat -> R^T [I | -t] (reference modules/geometry.py:106-108,147-163,207-213,249-255;
// three/quaternion.py:287-311,39-93) — and their autograd, as one forward and one analytic-VJP kernel, plus
// the batched Adam + ReduceLROnPlateau step over all hypotheses (reference pose/estimation.py:582-594,664-666).
#include "common.cuh"

namespace lf {

struct CamChain {
    float th, thc, s, c;     // |v|, clamped, sin(th)/thc, cos(th)
    float q0[4], n1, q1[4], n2, q[4];
    float R[9];
};

__device__ __forceinline__ void cam_chain(const float* v, CamChain& k) {
    k.th = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    k.thc = fmaxf(k.th, 1e-8f);
    k.c = cosf(k.th);
    k.s = 1.0f / k.thc * sinf(k.th);
    k.q0[0] = k.c; k.q0[1] = k.s * v[0]; k.q0[2] = k.s * v[1]; k.q0[3] = k.s * v[2];
    k.n1 = fmaxf(sqrtf(k.q0[0] * k.q0[0] + k.q0[1] * k.q0[1] + k.q0[2] * k.q0[2] + k.q0[3] * k.q0[3]), 1e-12f);
    for (int i = 0; i < 4; ++i) k.q1[i] = k.q0[i] / k.n1;
    k.n2 = fmaxf(sqrtf(k.q1[0] * k.q1[0] + k.q1[1] * k.q1[1] + k.q1[2] * k.q1[2] + k.q1[3] * k.q1[3]), 1e-12f);
    for (int i = 0; i < 4; ++i) k.q[i] = k.q1[i] / k.n2;
    const float w = k.q[0], x = k.q[1], y = k.q[2], z = k.q[3];
    const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
    k.R[0] = 1.f - (ty * y + tz * z); k.R[1] = ty * x - tz * w;         k.R[2] = tz * x + ty * w;
    k.R[3] = ty * x + tz * w;         k.R[4] = 1.f - (tx * x + tz * z); k.R[5] = tz * y - tx * w;
    k.R[6] = tz * x - ty * w;         k.R[7] = tz * y + tx * w;         k.R[8] = 1.f - (tx * x + ty * y);
}

// block layout: include/lfb200.h (object->camera)
__global__ void camera_o2c_fwd_kernel(const float* __restrict__ lq, const float* __restrict__ tr,
                                      const float* __restrict__ vp, const float* __restrict__ K,
                                      float* __restrict__ block, int n, float z_span, float half_cube) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    CamChain k;
    cam_chain(lq + 3 * i, k);
    const float* t = tr + 3 * i;
    float* b = block + (size_t)i * LF_CAM_STRIDE;
    for (int r = 0; r < 3; ++r) {
        // row r of cam_to_obj = R^T [I | -t]:  R^T[r][c] = R[c][r]
        const float a0 = k.R[0 * 3 + r], a1 = k.R[1 * 3 + r], a2 = k.R[2 * 3 + r];
        b[r * 4 + 0] = a0; b[r * 4 + 1] = a1; b[r * 4 + 2] = a2;
        b[r * 4 + 3] = a0 * (-t[0]) + a1 * (-t[1]) + a2 * (-t[2]);
    }
    const float* v = vp + 4 * i;
    b[12] = v[0]; b[13] = v[1]; b[14] = v[2] - v[0]; b[15] = v[3] - v[1];
    b[16] = K[12 * i + 2]; b[17] = K[12 * i + 6]; b[18] = K[12 * i + 0]; b[19] = K[12 * i + 5];
    b[20] = t[2] - z_span; b[21] = z_span; b[22] = half_cube;
    for (int j = 23; j < LF_CAM_STRIDE; ++j) b[j] = 0.f;
}

__device__ __forceinline__ void normalize_bwd(const float* q_out, float n, float n_raw_gt_eps, const float* g_out, float* g_in) {
    if (n_raw_gt_eps) {
        const float d = q_out[0] * g_out[0] + q_out[1] * g_out[1] + q_out[2] * g_out[2] + q_out[3] * g_out[3];
        for (int i = 0; i < 4; ++i) g_in[i] = (g_out[i] - q_out[i] * d) / n;
    } else {
        for (int i = 0; i < 4; ++i) g_in[i] = g_out[i] / n;
    }
}

// grad_block [n][LF_CAM_STRIDE] (only [0..15] and [20] are read) -> grads of the ten parameters
__global__ void camera_o2c_bwd_kernel(const float* __restrict__ lq, const float* __restrict__ tr,
                                      const float* __restrict__ gblock, float* __restrict__ g_lq,
                                      float* __restrict__ g_tr, float* __restrict__ g_vp, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    CamChain k;
    const float* v = lq + 3 * i;
    cam_chain(v, k);
    const float* t = tr + 3 * i;
    const float* gb = gblock + (size_t)i * LF_CAM_STRIDE;
    // GR[a][b] = G3[b][a] - gm[b] * t[a]   (M3[r][c] = R[c][r], m_r = -sum_k R[k][r] t[k])
    float GR[9], gt[3] = {0.f, 0.f, 0.f};
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) GR[a * 3 + b] = gb[b * 4 + a] - gb[b * 4 + 3] * t[a];
    for (int a = 0; a < 3; ++a) gt[a] = -(k.R[a * 3 + 0] * gb[3] + k.R[a * 3 + 1] * gb[7] + k.R[a * 3 + 2] * gb[11]);
    gt[2] += gb[20];                                                     // znear = t_z - z_span
    const float w = k.q[0], x = k.q[1], y = k.q[2], z = k.q[3];
    const float* G = GR;
    float gq[4];
    gq[0] = 2.f * (-z * G[1] + y * G[2] + z * G[3] - x * G[5] - y * G[6] + x * G[7]);
    gq[1] = 2.f * (y * G[1] + z * G[2] + y * G[3] - 2.f * x * G[4] - w * G[5] + z * G[6] + w * G[7] - 2.f * x * G[8]);
    gq[2] = 2.f * (-2.f * y * G[0] + x * G[1] + w * G[2] + x * G[3] + z * G[5] - w * G[6] + z * G[7] - 2.f * y * G[8]);
    gq[3] = 2.f * (-2.f * z * G[0] - w * G[1] + x * G[2] + w * G[3] - 2.f * z * G[4] + y * G[5] + x * G[6] + y * G[7]);
    float g1[4], g0[4];
    normalize_bwd(k.q, k.n2, k.n2 > 1e-12f, gq, g1);
    normalize_bwd(k.q1, k.n1, k.n1 > 1e-12f, g1, g0);
    // qexp: q0 = (cos th, s v), s = sin(th) / max(th, 1e-8)
    const float gs = g0[1] * v[0] + g0[2] * v[1] + g0[3] * v[2];
    const float dthc = k.th > 1e-8f ? 1.f : 0.f;
    const float ds_dth = k.c / k.thc - sinf(k.th) / (k.thc * k.thc) * dthc;
    const float gth = -sinf(k.th) * g0[0] + gs * ds_dth;
    for (int j = 0; j < 3; ++j) {
        const float dth_dv = k.th > 0.f ? v[j] / k.th : 0.f;
        g_lq[3 * i + j] = k.s * g0[1 + j] + gth * dth_dv;
        g_tr[3 * i + j] = gt[j];
    }
    // block viewport entries: x0, y0, w = x1 - x0, h = y1 - y0
    g_vp[4 * i + 0] = gb[12] - gb[14];
    g_vp[4 * i + 1] = gb[13] - gb[15];
    g_vp[4 * i + 2] = gb[14];
    g_vp[4 * i + 3] = gb[15];
}

// Adam (torch.optim.Adam single-tensor maths, betas (0.9, 0.999), eps 1e-8, no weight decay) with a per-row
// learning rate, followed by ReduceLROnPlateau(mode='min', threshold_mode='rel', cooldown 0, min_lr 0, eps 1e-8)
// on this iteration's ranking loss.  state: step (1 float), lr/best/num_bad [n].
__global__ void adam_plateau_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                    float* __restrict__ v, int n, int width, const float* __restrict__ step_count,
                                    const float* __restrict__ lr, float beta1, float beta2, float eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * width) return;
    const int row = i / width;
    const float step = step_count[0];
    const float bc1 = 1.f - powf(beta1, step);
    const float bc2s = sqrtf(1.f - powf(beta2, step));
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * (1.f - beta1);
    const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2s + eps;
    p[i] = p[i] - (lr[row] / bc1) * (mi / denom);
}

__global__ void plateau_kernel(const float* __restrict__ rank_loss, float* __restrict__ lr, float* __restrict__ best,
                               float* __restrict__ num_bad, int n, float threshold, float patience, float factor) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float metric = rank_loss[i];
    float nb = num_bad[i];
    if (metric < best[i] * (1.f - threshold)) { best[i] = metric; nb = 0.f; }
    else nb += 1.f;
    if (nb > patience) {
        const float new_lr = lr[i] * factor;
        if (lr[i] - new_lr > 1e-8f) lr[i] = new_lr;
        nb = 0.f;
    }
    num_bad[i] = nb;
}

// One refinement iteration's loss combination and bookkeeping (reference pose/estimation.py:611-660): the ranking and
// optimisation losses as the weighted sums of the terms (left to right, like Python's sum() over the loss dict), the
// gradient of mean_n(optim) w.r.t. the terms (what optim.mean().backward() hands the loss head: w_opt[k] / N), and the
// snapshot of this iteration into the chunk history at `slot`, which then advances.  One CTA.
__global__ void refine_record_kernel(const float* __restrict__ terms, int n, int k, const float* __restrict__ w_rank,
                                     const float* __restrict__ w_opt, const float* __restrict__ lq,
                                     const float* __restrict__ tr, float* __restrict__ rank, float* __restrict__ gterms,
                                     float* __restrict__ h_rank, float* __restrict__ h_optim, float* __restrict__ h_terms,
                                     float* __restrict__ h_lq, float* __restrict__ h_tr, long long* __restrict__ slot,
                                     int chunk, float* __restrict__ step_count) {
    const long long sl = *slot;
    const float inv_n = 1.f / (float)n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float r = 0.f, o = 0.f;
        for (int j = 0; j < k; ++j) {
            const float t = terms[i * k + j];
            r = __fadd_rn(r, __fmul_rn(w_rank[j], t));
            o = __fadd_rn(o, __fmul_rn(w_opt[j], t));
            gterms[i * k + j] = __fmul_rn(inv_n, w_opt[j]);
            h_terms[(sl * k + j) * n + i] = t;
        }
        rank[i] = r;
        h_rank[sl * n + i] = r;
        h_optim[sl * n + i] = o;
        for (int c = 0; c < 3; ++c) {
            h_lq[(sl * n + i) * 3 + c] = lq[i * 3 + c];
            h_tr[(sl * n + i) * 3 + c] = tr[i * 3 + c];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *slot = (sl + 1) % chunk;
        if (step_count) *step_count += 1.f;
    }
}

}  // namespace lf

using namespace lf;

extern "C" int lf_refine_record(const float* terms, int n, int k, const float* w_rank, const float* w_opt,
                                const float* log_quaternion, const float* translation, float* rank, float* grad_terms,
                                float* h_rank, float* h_optim, float* h_terms, float* h_lq, float* h_tr,
                                long long* slot, int chunk, float* step_count, void* stream) {
    LF_CHECK_ARG(terms && w_rank && w_opt && log_quaternion && translation && rank && grad_terms && h_rank && h_optim &&
                 h_terms && h_lq && h_tr && slot && n > 0 && k > 0 && chunk > 0, "refine_record: bad arguments");
    refine_record_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(terms, n, k, w_rank, w_opt, log_quaternion, translation, rank,
                                                              grad_terms, h_rank, h_optim, h_terms, h_lq, h_tr, slot, chunk,
                                                              step_count);
    LF_RETURN_LAUNCH();
}

extern "C" int lf_camera_o2c_fwd(const float* log_quaternion, const float* translation, const float* viewport,
                                 const float* intrinsic, float* block, int n, float z_span, float cube_size,
                                 void* stream) {
    LF_CHECK_ARG(log_quaternion && translation && viewport && intrinsic && block && n > 0, "camera_o2c_fwd: bad arguments");
    camera_o2c_fwd_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(log_quaternion, translation, viewport,
                                                                          intrinsic, block, n, z_span, cube_size * 0.5f);
    LF_RETURN_LAUNCH();
}

extern "C" int lf_camera_o2c_bwd(const float* log_quaternion, const float* translation, const float* grad_block,
                                 float* grad_log_quaternion, float* grad_translation, float* grad_viewport, int n,
                                 void* stream) {
    LF_CHECK_ARG(log_quaternion && translation && grad_block && grad_log_quaternion && grad_translation &&
                 grad_viewport && n > 0, "camera_o2c_bwd: bad arguments");
    camera_o2c_bwd_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(log_quaternion, translation, grad_block,
                                                                          grad_log_quaternion, grad_translation,
                                                                          grad_viewport, n);
    LF_RETURN_LAUNCH();
}

extern "C" int lf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int n, int width,
                            const float* step_count, const float* lr, float beta1, float beta2, float eps,
                            void* stream) {
    LF_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && step_count && lr && n > 0 && width > 0, "adam_step: bad arguments");
    adam_plateau_kernel<<<(n * width + 127) / 128, 128, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n,
                                                                                   width, step_count, lr, beta1, beta2, eps);
    LF_RETURN_LAUNCH();
}

extern "C" int lf_plateau_step(const float* rank_loss, float* lr, float* best, float* num_bad, int n, float threshold,
                               float patience, float factor, void* stream) {
    LF_CHECK_ARG(rank_loss && lr && best && num_bad && n > 0, "plateau_step: bad arguments");
    plateau_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(rank_loss, lr, best, num_bad, n, threshold, patience, factor);
    LF_RETURN_LAUNCH();
}
