/* TEST INFRASTRUCTURE — independent plain-C restatement (double precision) of the two voxel
 * resamplers of the LatentFusion hot path, so that a kernel bug and a harness bug cannot cancel:
 *
 *   camera chain   latentfusion/modules/geometry.py:106-108,147-163,207-213,249-255
 *                  latentfusion/three/quaternion.py:287-311 (qexp), :39-93 (quat_to_mat)
 *   object->camera latentfusion/modules/geometry.py:469-531, :669-690
 *   camera->object latentfusion/modules/geometry.py:599-611, :625-657
 *   sampler        ATen/native/GridSampler.h semantics (torch is a third-party dependency of the
 *                  reference, not vendored): grid_sampler_unnormalize (align_corners=False),
 *                  clip_coordinates (padding_mode='border'), trilinear weights from floor, corner
 *                  reads outside the volume contribute 0.
 *
 * Layout is the reference's: volumes [n][C][S][S][S] (channel-first), fp32 in, fp64 arithmetic.
 * Built by __graft_entry__.build() into oracle/libresample_ref.so; used only by tests/. */
#include <math.h>
#include <stddef.h>

typedef struct {
    double log_q[3], trans[3], viewport[4];   /* the ten learnable floats */
    double K[12];                              /* 3x4 intrinsic, row major */
    double z_span, cube;
} ref_camera;

static void rotation_from_logq(const double v[3], double R[9]) {
    /* qexp of a pure quaternion, then F.normalize twice (eps 1e-12), then quat_to_mat */
    double theta = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double tc = theta < 1e-8 ? 1e-8 : theta;
    double s = 1.0 / tc * sin(theta);
    double q[4] = {cos(theta), s * v[0], s * v[1], s * v[2]};
    for (int rep = 0; rep < 2; ++rep) {
        double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        if (nrm < 1e-12) nrm = 1e-12;
        for (int i = 0; i < 4; ++i) q[i] /= nrm;
    }
    double w = q[0], x = q[1], y = q[2], z = q[3];
    double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    R[0] = 1 - (ty * y + tz * z); R[1] = ty * x - tz * w;       R[2] = tz * x + ty * w;
    R[3] = ty * x + tz * w;       R[4] = 1 - (tx * x + tz * z); R[5] = tz * y - tx * w;
    R[6] = tz * x - ty * w;       R[7] = tz * y + tx * w;       R[8] = 1 - (tx * x + ty * y);
}

/* trilinear sample of one channel plane vol[S][S][S] at normalised grid coords (gx -> W, gy -> H, gz -> D) */
static double sample(const float* vol, int S, double gx, double gy, double gz) {
    double c[3] = {gx, gy, gz};
    double ic[3];
    for (int a = 0; a < 3; ++a) {
        double v = ((c[a] + 1.0) * S - 1.0) / 2.0;
        if (v < 0) v = 0;
        if (v > S - 1) v = S - 1;
        ic[a] = v;
    }
    int x0 = (int)floor(ic[0]), y0 = (int)floor(ic[1]), z0 = (int)floor(ic[2]);
    double fx = ic[0] - x0, fy = ic[1] - y0, fz = ic[2] - z0;
    double acc = 0.0;
    for (int dz = 0; dz < 2; ++dz)
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) {
                int x = x0 + dx, y = y0 + dy, z = z0 + dz;
                if (x >= S || y >= S || z >= S) continue;          /* out of bounds -> 0 */
                double w = (dx ? fx : 1 - fx) * (dy ? fy : 1 - fy) * (dz ? fz : 1 - fz);
                acc += w * (double)vol[((size_t)z * S + y) * S + x];
            }
    return acc;
}

static double linspace(double a, double b, int n, int i) { return a + (b - a) * (double)i / (double)(n - 1); }

/* ObjectToCameraTransform: vol [C][S][S][S] (one cube), cams [n] -> out [n][C][S][S][S] (double) */
void lf_ref_object_to_camera(const float* vol, const ref_camera* cams, int n, int C, int S, double* out) {
    const size_t S3 = (size_t)S * S * S;
    for (int cidx = 0; cidx < n; ++cidx) {
        const ref_camera* cam = &cams[cidx];
        double R[9];
        rotation_from_logq(cam->log_q, R);
        const double vw = cam->viewport[2] - cam->viewport[0], vh = cam->viewport[3] - cam->viewport[1];
        const double znear = cam->trans[2] - cam->z_span;
        const double u0 = cam->K[2], v0 = cam->K[6], fu = cam->K[0], fv = cam->K[5];
        for (int k = 0; k < S; ++k)
            for (int j = 0; j < S; ++j)
                for (int i = 0; i < S; ++i) {
                    /* frustum point: u,v over the viewport, z over [znear, znear + z_span] (near half only) */
                    double u = linspace(0, 1, S, i) * vw + cam->viewport[0];
                    double v = linspace(0, 1, S, j) * vh + cam->viewport[1];
                    double z = linspace(0, 1, S, k) * cam->z_span + znear;
                    double x = (u - u0) / fu * z, y = (v - v0) / fv * z;
                    /* cam_to_obj = R^T * T^-1 : p_obj = R^T (p_cam - t) */
                    double px = x - cam->trans[0], py = y - cam->trans[1], pz = z - cam->trans[2];
                    double ox = R[0] * px + R[3] * py + R[6] * pz;
                    double oy = R[1] * px + R[4] * py + R[7] * pz;
                    double oz = R[2] * px + R[5] * py + R[8] * pz;
                    double half = cam->cube / 2.0;
                    for (int c = 0; c < C; ++c)
                        out[((size_t)cidx * C + c) * S3 + ((size_t)k * S + j) * S + i] =
                            sample(vol + (size_t)c * S3, S, ox / half, oy / half, oz / half);
                }
    }
}

/* CameraToObjectTransform: vol [n][C][S][S][S], cams [n] -> out [n][C][S][S][S] (double) */
void lf_ref_camera_to_object(const float* vol, const ref_camera* cams, int n, int C, int S, double* out) {
    const size_t S3 = (size_t)S * S * S;
    for (int cidx = 0; cidx < n; ++cidx) {
        const ref_camera* cam = &cams[cidx];
        double R[9];
        rotation_from_logq(cam->log_q, R);
        const double vw = cam->viewport[2] - cam->viewport[0], vh = cam->viewport[3] - cam->viewport[1];
        const double znear = cam->trans[2] - cam->z_span, zfar = cam->trans[2] + cam->z_span;
        for (int k = 0; k < S; ++k)
            for (int j = 0; j < S; ++j)
                for (int i = 0; i < S; ++i) {
                    double half = cam->cube / 2.0;
                    double x = linspace(-half, half, S, i), y = linspace(-half, half, S, j), z = linspace(-half, half, S, k);
                    /* obj_to_cam = T * R */
                    double cx = R[0] * x + R[1] * y + R[2] * z + cam->trans[0];
                    double cy = R[3] * x + R[4] * y + R[5] * z + cam->trans[1];
                    double cz = R[6] * x + R[7] * y + R[8] * z + cam->trans[2];
                    const double* K = cam->K;
                    double p0 = K[0] * cx + K[1] * cy + K[2] * cz + K[3];
                    double p1 = K[4] * cx + K[5] * cy + K[6] * cz + K[7];
                    double p2 = K[8] * cx + K[9] * cy + K[10] * cz + K[11];
                    double gx = ((p0 / p2 - cam->viewport[0]) / vw) * 2 - 1;
                    double gy = ((p1 / p2 - cam->viewport[1]) / vh) * 2 - 1;
                    double gz = (p2 - znear) / (zfar - znear);            /* [0,1]: the reference's quirk */
                    for (int c = 0; c < C; ++c)
                        out[((size_t)cidx * C + c) * S3 + ((size_t)k * S + j) * S + i] =
                            sample(vol + ((size_t)cidx * C + c) * S3, S, gx, gy, gz);
                }
    }
}
