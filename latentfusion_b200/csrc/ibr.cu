// IBR colour branch: image-based rendering of the reference views into the output views.
//
//   reproject   latentfusion/ibr.py:11-93   depth_to_warp_field + reproject_views
//   blend       latentfusion/ibr.py:223-224, :231-234   weighted sum over the input views
//   warp_blend  latentfusion/ibr.py:237-249  warp_blend_logits
//
// The reference builds, per (output view, input view) pair, a [H*W, 3] point list, two batched matrix
// products, a [V_o*V_i, H, W, 2] sampling grid and two F.grid_sample calls (bilinear, zeros padding,
// align_corners=False), plus a [V_o*V_i, 1, H, W] transformed-depth image that exists only to be sampled.
// Here one thread owns one output pixel of one (o, i) pair: the warp coordinate is generated in registers from
// two 48-float camera blocks, the four taps of the colour image are gathered directly, and the transformed
// depth is evaluated at the four taps on the fly (it is a closed form of depth_in at the tap).  The reprojection is
// forward only (the recon networks are frozen in tools/train/train_ibr.py: _render_reprojections runs without grad);
// the two blend heads have backward kernels to the logits / weights that the IBR generator is trained through
// (train_ibr.py:367-376).
#include "common.cuh"

namespace lf {

// IBR camera block (LF_IBR_CAM_STRIDE floats):
//   [0,12)  cam_to_obj rows 0-2     [12,24) obj_to_cam rows 0-2     [24,36) obj_to_image = K * obj_to_cam
//   [36,40) viewport x0, y0, width, height      [40,44) u0, v0, fu, fv      [44] znear - eps   [45] zfar + eps
struct Tap {
    int x0, y0;
    float w[4];          // nw, ne, sw, se (ATen order)
    bool ok[4];
};

// ATen GridSampler.h, bilinear / zeros / align_corners=False
__device__ __forceinline__ Tap make_tap(float gx, float gy, int W, int H) {
    Tap t;
    const float ix = ((gx + 1.f) * (float)W - 1.f) / 2.f;
    const float iy = ((gy + 1.f) * (float)H - 1.f) / 2.f;
    const bool finite = (ix > -2.f) && (ix < (float)W + 1.f) && (iy > -2.f) && (iy < (float)H + 1.f);   // false for NaN too
    const float fx = finite ? floorf(ix) : -2.f, fy = finite ? floorf(iy) : -2.f;
    t.x0 = (int)fx; t.y0 = (int)fy;
    const float x1 = fx + 1.f, y1 = fy + 1.f;
    t.w[0] = (x1 - ix) * (y1 - iy);
    t.w[1] = (ix - fx) * (y1 - iy);
    t.w[2] = (x1 - ix) * (iy - fy);
    t.w[3] = (ix - fx) * (iy - fy);
    const bool xa = t.x0 >= 0 && t.x0 < W, xb = t.x0 + 1 >= 0 && t.x0 + 1 < W;
    const bool ya = t.y0 >= 0 && t.y0 < H, yb = t.y0 + 1 >= 0 && t.y0 + 1 < H;
    t.ok[0] = finite && xa && ya; t.ok[1] = finite && xb && ya; t.ok[2] = finite && xa && yb; t.ok[3] = finite && xb && yb;
    return t;
}

__device__ __forceinline__ float tap_sum(const Tap& t, const float* __restrict__ plane, int W) {
    float acc = 0.f;
    if (t.ok[0]) acc += __ldg(plane + t.y0 * W + t.x0) * t.w[0];
    if (t.ok[1]) acc += __ldg(plane + t.y0 * W + t.x0 + 1) * t.w[1];
    if (t.ok[2]) acc += __ldg(plane + (t.y0 + 1) * W + t.x0) * t.w[2];
    if (t.ok[3]) acc += __ldg(plane + (t.y0 + 1) * W + t.x0 + 1) * t.w[3];
    return acc;
}

__global__ void __launch_bounds__(256)
ibr_reproject_kernel(const float* __restrict__ image_in, const float* __restrict__ depth_in,
                     const float* __restrict__ depth_out, const float* __restrict__ cam_out,
                     const float* __restrict__ cam_in, float* __restrict__ image_reproj,
                     float* __restrict__ depth_reproj, int VO, int VI, int C, int H, int W) {
    __shared__ float co[LF_IBR_CAM_STRIDE], ci[LF_IBR_CAM_STRIDE];
    const int pair = blockIdx.y;                     // o * VI + i
    const int o = pair / VI, i = pair - o * VI;
    if (threadIdx.x < LF_IBR_CAM_STRIDE) {
        co[threadIdx.x] = cam_out[o * LF_IBR_CAM_STRIDE + threadIdx.x];
        ci[threadIdx.x] = cam_in[i * LF_IBR_CAM_STRIDE + threadIdx.x];
    }
    __syncthreads();
    const int HW = H * W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int y = p / W, x = p - y * W;

    // output pixel -> camera-o coordinates at the denormalised output depth (geometry.py:533-545, :555-558)
    const float d = (depth_out[(int64_t)o * HW + p] / 2.f + 0.5f) * (co[45] - co[44]) + co[44];
    const float u = linspace_at(0.f, 1.f, W, x) * co[38] + co[36];
    const float v = linspace_at(0.f, 1.f, H, y) * co[39] + co[37];
    const float xc = (u - co[40]) / co[42] * d;
    const float yc = (v - co[41]) / co[43] * d;
    // -> object -> pixel of input view i (ibr.py:28-37)
    const float ox = co[0] * xc + co[1] * yc + co[2] * d + co[3];
    const float oy = co[4] * xc + co[5] * yc + co[6] * d + co[7];
    const float oz = co[8] * xc + co[9] * yc + co[10] * d + co[11];
    const float p0 = ci[24] * ox + ci[25] * oy + ci[26] * oz + ci[27];
    const float p1 = ci[28] * ox + ci[29] * oy + ci[30] * oz + ci[31];
    const float p2 = ci[32] * ox + ci[33] * oy + ci[34] * oz + ci[35];
    const float gx = ((p0 / p2 - ci[36]) / ci[38]) * 2.f - 1.f;
    const float gy = ((p1 / p2 - ci[37]) / ci[39]) * 2.f - 1.f;
    const Tap t = make_tap(gx, gy, W, H);

    // colour: C planes of input view i
    const float* img = image_in + (int64_t)i * C * HW;
    float* outp = image_reproj + (int64_t)pair * C * HW + p;
    for (int c = 0; c < C; ++c) outp[(int64_t)c * HW] = tap_sum(t, img + (int64_t)c * HW, W);

    // depth of input view i, carried into output camera o, evaluated at the four taps (ibr.py:72-86).
    // NB: depth_in is used as given (the reference does not denormalise it here).
    const float lo = co[44], hi = co[45];
    float dacc = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (!t.ok[q]) continue;
        const int xx = t.x0 + (q & 1), yy = t.y0 + (q >> 1);
        const float din = __ldg(depth_in + (int64_t)i * HW + yy * W + xx);
        const float ui = linspace_at(0.f, 1.f, W, xx) * ci[38] + ci[36];
        const float vi = linspace_at(0.f, 1.f, H, yy) * ci[39] + ci[37];
        const float xi = (ui - ci[40]) / ci[42] * din;
        const float yi = (vi - ci[41]) / ci[43] * din;
        const float bx = ci[0] * xi + ci[1] * yi + ci[2] * din + ci[3];
        const float by = ci[4] * xi + ci[5] * yi + ci[6] * din + ci[7];
        const float bz = ci[8] * xi + ci[9] * yi + ci[10] * din + ci[11];
        const float z = co[20] * bx + co[21] * by + co[22] * bz + co[23];
        float nd = (z - lo) / (hi - lo);
        nd = fminf(fmaxf(nd, 0.f), 1.f) * 2.f - 1.f;
        dacc += nd * t.w[q];
    }
    depth_reproj[(int64_t)pair * HW + p] = dacc;
}

// out[b][c][p] = sum_i wts[b][i][per_pixel ? p : 0] * img[b][i][c][p]      (ibr.py:223-224, :231-234)
__global__ void __launch_bounds__(256)
ibr_blend_kernel(const float* __restrict__ img, const float* __restrict__ wts, float* __restrict__ out,
                 int B, int VI, int C, int HW, int per_pixel) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)B * C * HW) return;
    const int p = (int)(e % HW);
    const int c = (int)((e / HW) % C);
    const int b = (int)(e / ((int64_t)HW * C));
    float acc = 0.f;
    for (int i = 0; i < VI; ++i) {
        const float w = per_pixel ? __ldg(wts + ((int64_t)b * VI + i) * HW + p) : __ldg(wts + (int64_t)b * VI + i);
        acc += w * __ldg(img + (((int64_t)b * VI + i) * C + c) * HW + p);
    }
    out[e] = acc;
}

// ibr.py:237-249: softmax over the views of the blend logits, bounded flow refinement of each reprojection,
// weighted sum.  One thread per output pixel of one batch element.
__global__ void __launch_bounds__(256)
ibr_warp_blend_kernel(const float* __restrict__ logits, const float* __restrict__ image_reproj, float flow_size,
                      float* __restrict__ image, float* __restrict__ weights, float* __restrict__ flow_dx,
                      float* __restrict__ flow_dy, int B, int VI, int C, int H, int W) {
    const int HW = H * W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (p >= HW) return;
    const int y = p / W, x = p - y * W;
    const float* lg = logits + (int64_t)b * 3 * VI * HW + p;
    float m = -INFINITY;
    for (int i = 0; i < VI; ++i) m = fmaxf(m, lg[(int64_t)i * HW]);
    float den = 0.f;
    for (int i = 0; i < VI; ++i) den += expf(lg[(int64_t)i * HW] - m);
    const float bx = linspace_at(-1.f, 1.f, W, x), by = linspace_at(-1.f, 1.f, H, y);
    float acc[8];
    for (int c = 0; c < C && c < 8; ++c) acc[c] = 0.f;
    for (int i = 0; i < VI; ++i) {
        const float wgt = expf(lg[(int64_t)i * HW] - m) / den;
        const float dx = flow_size / (float)W * tanhf(lg[(int64_t)(VI + i) * HW]);
        const float dy = flow_size / (float)H * tanhf(lg[(int64_t)(2 * VI + i) * HW]);
        const int64_t wi = ((int64_t)b * VI + i) * HW + p;
        weights[wi] = wgt; flow_dx[wi] = dx; flow_dy[wi] = dy;
        const float gx = fminf(fmaxf(bx + dx, -1.f), 1.f), gy = fminf(fmaxf(by + dy, -1.f), 1.f);
        const Tap t = make_tap(gx, gy, W, H);
        const float* src = image_reproj + ((int64_t)b * VI + i) * C * HW;
        for (int c = 0; c < C && c < 8; ++c) acc[c] += wgt * tap_sum(t, src + (int64_t)c * HW, W);
    }
    for (int c = 0; c < C && c < 8; ++c) image[((int64_t)b * C + c) * HW + p] = acc[c];
}

// ---- backward of the blend heads (tools/train/train_ibr.py:367-376 trains the generator through them) ----
// d/d weights of out[b][c][p] = sum_i w[b][i][p] * img[b][i][c][p]:  gw[b][i][p] = sum_c g[b][c][p] * img[b][i][c][p]
__global__ void __launch_bounds__(256)
ibr_blend_bwd_kernel(const float* __restrict__ g, const float* __restrict__ img, float* __restrict__ gw,
                     int B, int VI, int C, int HW) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)B * VI * HW) return;
    const int p = (int)(e % HW);
    const int i = (int)((e / HW) % VI);
    const int b = (int)(e / ((int64_t)HW * VI));
    float acc = 0.f;
    for (int c = 0; c < C; ++c)
        acc += __ldg(g + ((int64_t)b * C + c) * HW + p) * __ldg(img + (((int64_t)b * VI + i) * C + c) * HW + p);
    gw[e] = acc;
}

// bilinear sample of one plane and its derivatives w.r.t. the unnormalised coordinates (ATen grid_sampler_2d_backward,
// zeros padding: taps outside the image read as 0 and still carry weight derivatives)
__device__ __forceinline__ void tap_grad(float gx, float gy, int W, int H, const float* __restrict__ plane,
                                         float& val, float& dix, float& diy) {
    const float ix = ((gx + 1.f) * (float)W - 1.f) / 2.f;
    const float iy = ((gy + 1.f) * (float)H - 1.f) / 2.f;
    val = dix = diy = 0.f;
    if (!((ix > -2.f) && (ix < (float)W + 1.f) && (iy > -2.f) && (iy < (float)H + 1.f))) return;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float ax = (fx + 1.f) - ix, bx = ix - fx, ay = (fy + 1.f) - iy, by = iy - fy;
    const bool xa = x0 >= 0 && x0 < W, xb = x0 + 1 >= 0 && x0 + 1 < W, ya = y0 >= 0 && y0 < H, yb = y0 + 1 >= 0 && y0 + 1 < H;
    const float nw = (xa && ya) ? __ldg(plane + y0 * W + x0) : 0.f;
    const float ne = (xb && ya) ? __ldg(plane + y0 * W + x0 + 1) : 0.f;
    const float sw = (xa && yb) ? __ldg(plane + (y0 + 1) * W + x0) : 0.f;
    const float se = (xb && yb) ? __ldg(plane + (y0 + 1) * W + x0 + 1) : 0.f;
    val = nw * (ax * ay) + ne * (bx * ay) + sw * (ax * by) + se * (bx * by);
    dix = (ne - nw) * ay + (se - sw) * by;
    diy = (sw - nw) * ax + (se - ne) * bx;
}

// backward of warp_blend_logits to the logits [B][3*Vi][H][W]; upstream gradients of all four outputs
// (g_w / g_dx / g_dy may be null)
__global__ void __launch_bounds__(256)
ibr_warp_blend_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ image_reproj, float flow_size,
                          const float* __restrict__ g_image, const float* __restrict__ g_w, const float* __restrict__ g_dx,
                          const float* __restrict__ g_dy, float* __restrict__ g_logits, int B, int VI, int C, int H, int W) {
    const int HW = H * W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (p >= HW) return;
    const int y = p / W, x = p - y * W;
    const float* lg = logits + (int64_t)b * 3 * VI * HW + p;
    float* gl = g_logits + (int64_t)b * 3 * VI * HW + p;
    float m = -INFINITY;
    for (int i = 0; i < VI; ++i) m = fmaxf(m, lg[(int64_t)i * HW]);
    float den = 0.f;
    for (int i = 0; i < VI; ++i) den += expf(lg[(int64_t)i * HW] - m);
    const float bx = linspace_at(-1.f, 1.f, W, x), by = linspace_at(-1.f, 1.f, H, y);
    float gc[8];
    for (int c = 0; c < C && c < 8; ++c) gc[c] = __ldg(g_image + ((int64_t)b * C + c) * HW + p);
    // pass 1: a_i = <g, S_i>, the flow gradients; A = sum_i w_i (a_i + gw_i)
    float A = 0.f;
    for (int i = 0; i < VI; ++i) {
        const float wgt = expf(lg[(int64_t)i * HW] - m) / den;
        const float tx = tanhf(lg[(int64_t)(VI + i) * HW]), ty = tanhf(lg[(int64_t)(2 * VI + i) * HW]);
        const float ux = bx + flow_size / (float)W * tx, uy = by + flow_size / (float)H * ty;
        const float gx = fminf(fmaxf(ux, -1.f), 1.f), gy = fminf(fmaxf(uy, -1.f), 1.f);
        const float* src = image_reproj + ((int64_t)b * VI + i) * C * HW;
        float a = 0.f, ddx = 0.f, ddy = 0.f;
        for (int c = 0; c < C && c < 8; ++c) {
            float v, dix, diy;
            tap_grad(gx, gy, W, H, src + (int64_t)c * HW, v, dix, diy);
            a += gc[c] * v; ddx += gc[c] * dix; ddy += gc[c] * diy;
        }
        const int64_t wi = ((int64_t)b * VI + i) * HW + p;
        const float up = a + (g_w != nullptr ? __ldg(g_w + wi) : 0.f);
        A += wgt * up;
        gl[(int64_t)i * HW] = up;                               // finished in pass 2
        // clamp passes the gradient inside [-1, 1] (inclusive, as torch.clamp); unnormalisation d ix / d gx = W / 2
        float gfx = (ux >= -1.f && ux <= 1.f) ? wgt * ddx * (0.5f * (float)W) : 0.f;
        float gfy = (uy >= -1.f && uy <= 1.f) ? wgt * ddy * (0.5f * (float)H) : 0.f;
        if (g_dx != nullptr) gfx += __ldg(g_dx + wi);
        if (g_dy != nullptr) gfy += __ldg(g_dy + wi);
        gl[(int64_t)(VI + i) * HW] = gfx * (flow_size / (float)W) * (1.f - tx * tx);
        gl[(int64_t)(2 * VI + i) * HW] = gfy * (flow_size / (float)H) * (1.f - ty * ty);
    }
    for (int i = 0; i < VI; ++i) {                              // softmax backward: w_i (up_i - A)
        const float wgt = expf(lg[(int64_t)i * HW] - m) / den;
        gl[(int64_t)i * HW] = wgt * (gl[(int64_t)i * HW] - A);
    }
}

}  // namespace lf

using namespace lf;

extern "C" int lf_ibr_reproject_fwd(const float* image_in, const float* depth_in, const float* depth_out,
                                    const float* cam_out, const float* cam_in, float* image_reproj,
                                    float* depth_reproj, int vo, int vi, int c, int h, int w, void* stream) {
    LF_CHECK_ARG(image_in && depth_in && depth_out && cam_out && cam_in && image_reproj && depth_reproj, "ibr_reproject: null pointer");
    LF_CHECK_ARG(vo > 0 && vi > 0 && c > 0 && h > 1 && w > 1, "ibr_reproject: bad extents");
    LF_CHECK_ARG((int64_t)vo * vi < 65536 && (int64_t)h * w < (1ll << 30), "ibr_reproject: too many view pairs / pixels");
    dim3 grid((unsigned)((h * w + 255) / 256), (unsigned)(vo * vi));
    ibr_reproject_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(image_in, depth_in, depth_out, cam_out, cam_in,
                                                                  image_reproj, depth_reproj, vo, vi, c, h, w);
    LF_RETURN_LAUNCH();
}

extern "C" int lf_ibr_blend_fwd(const float* img, const float* wts, float* out, int b, int vi, int c, int hw,
                                int per_pixel, void* stream) {
    LF_CHECK_ARG(img && wts && out, "ibr_blend: null pointer");
    LF_CHECK_ARG(b > 0 && vi > 0 && c > 0 && hw > 0, "ibr_blend: bad extents");
    const int64_t total = (int64_t)b * c * hw;
    LF_CHECK_ARG((total + 255) / 256 < (1ll << 31), "ibr_blend: too many elements");
    ibr_blend_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(img, wts, out, b, vi, c, hw, per_pixel);
    LF_RETURN_LAUNCH();
}

extern "C" int lf_ibr_warp_blend_fwd(const float* logits, const float* image_reproj, float flow_size, float* image,
                                     float* weights, float* flow_dx, float* flow_dy, int b, int vi, int c, int h, int w,
                                     void* stream) {
    LF_CHECK_ARG(logits && image_reproj && image && weights && flow_dx && flow_dy, "ibr_warp_blend: null pointer");
    LF_CHECK_ARG(b > 0 && b < 65536 && vi > 0 && c > 0 && c <= 8 && h > 1 && w > 1, "ibr_warp_blend: bad extents (C <= 8)");
    dim3 grid((unsigned)((h * w + 255) / 256), (unsigned)b);
    ibr_warp_blend_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, image_reproj, flow_size, image, weights,
                                                                   flow_dx, flow_dy, b, vi, c, h, w);
    LF_RETURN_LAUNCH();
}

extern "C" int lf_ibr_blend_bwd(const float* grad_out, const float* img, float* grad_wts, int b, int vi, int c, int hw,
                                void* stream) {
    LF_CHECK_ARG(grad_out && img && grad_wts, "ibr_blend_bwd: null pointer");
    LF_CHECK_ARG(b > 0 && vi > 0 && c > 0 && hw > 0, "ibr_blend_bwd: bad extents");
    const int64_t total = (int64_t)b * vi * hw;
    LF_CHECK_ARG((total + 255) / 256 < (1ll << 31), "ibr_blend_bwd: too many elements");
    ibr_blend_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(grad_out, img, grad_wts, b, vi, c, hw);
    LF_RETURN_LAUNCH();
}

extern "C" int lf_ibr_warp_blend_bwd(const float* logits, const float* image_reproj, float flow_size, const float* grad_image,
                                     const float* grad_weights, const float* grad_flow_dx, const float* grad_flow_dy,
                                     float* grad_logits, int b, int vi, int c, int h, int w, void* stream) {
    LF_CHECK_ARG(logits && image_reproj && grad_image && grad_logits, "ibr_warp_blend_bwd: null pointer");
    LF_CHECK_ARG(b > 0 && b < 65536 && vi > 0 && c > 0 && c <= 8 && h > 1 && w > 1, "ibr_warp_blend_bwd: bad extents (C <= 8)");
    dim3 grid((unsigned)((h * w + 255) / 256), (unsigned)b);
    ibr_warp_blend_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, image_reproj, flow_size, grad_image, grad_weights,
                                                                       grad_flow_dx, grad_flow_dy, grad_logits, b, vi, c, h, w);
    LF_RETURN_LAUNCH();
}
