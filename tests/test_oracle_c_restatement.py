"""CPU: the plain-C double-precision restatement of the two resamplers (oracle/resample_ref.c) against
(a) golden vectors from the unmodified reference and (b) the PyTorch oracle, plus a finite-difference check
of the camera gradients the reference's autograd produced (guards the border-clamp gradient rule)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from tests import parity_helpers as ph

ROOT = ph.ROOT


class RefCamera(ctypes.Structure):
    _fields_ = [('log_q', ctypes.c_double * 3), ('trans', ctypes.c_double * 3), ('viewport', ctypes.c_double * 4),
                ('K', ctypes.c_double * 12), ('z_span', ctypes.c_double), ('cube', ctypes.c_double)]


@pytest.fixture(scope='module')
def cref():
    src = os.path.join(ROOT, 'oracle', 'resample_ref.c')
    so = os.path.join(ROOT, 'oracle', 'libresample_ref.so')
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['gcc', '-O2', '-shared', '-fPIC', '-o', so, src, '-lm'])
    return ctypes.CDLL(so)


def cams_struct(d, lq=None, tr=None, vp=None):
    n = d['intrinsic'].shape[0]
    arr = (RefCamera * n)()
    lq = d['log_quaternion'].double() if lq is None else lq
    tr = d['translation'].double() if tr is None else tr
    vp = d['viewport'].double() if vp is None else vp
    for i in range(n):
        arr[i].log_q[:] = lq[i].tolist()
        arr[i].trans[:] = tr[i].tolist()
        arr[i].viewport[:] = vp[i].tolist()
        arr[i].K[:] = d['intrinsic'][i].double().reshape(-1).tolist()
        arr[i].z_span, arr[i].cube = 0.5, 1.0
    return arr, n


def run_o2c(lib, vol, cams, n):
    C, S = vol.shape[1], vol.shape[-1]
    out = np.zeros((n, C, S, S, S), dtype=np.float64)
    v = np.ascontiguousarray(vol[0].numpy(), dtype=np.float32)
    lib.lf_ref_object_to_camera(v.ctypes.data_as(ctypes.c_void_p), cams, n, C, S, out.ctypes.data_as(ctypes.c_void_p))
    return torch.from_numpy(out)


def test_c_o2c_matches_reference_golden(cref):
    g = ph.Golden()
    cams, n = cams_struct(g.cam('hyp_cam'))
    out = run_o2c(cref, g['o2c.vol'], cams, n)
    torch.testing.assert_close(out.float(), g['o2c.out'], atol=1e-4, rtol=1e-3)


def test_c_c2o_matches_reference_golden(cref):
    g = ph.Golden()
    cams, n = cams_struct(g.cam('ref_cam'))
    vol = g['c2o.vol']
    V, C, S = vol.shape[0], vol.shape[1], vol.shape[-1]
    out = np.zeros((V, C, S, S, S), dtype=np.float64)
    v = np.ascontiguousarray(vol.numpy(), dtype=np.float32)
    cref.lf_ref_camera_to_object(v.ctypes.data_as(ctypes.c_void_p), cams, n, C, S, out.ctypes.data_as(ctypes.c_void_p))
    torch.testing.assert_close(torch.from_numpy(out).float(), g['c2o.out'], atol=1e-4, rtol=1e-3)


def test_camera_gradients_by_finite_differences(cref):
    """d/d(log_quaternion, translation, viewport) of sum(out * w): central differences of the C restatement
    (fp64) vs the gradients the reference's autograd produced (golden).  White-noise volumes make the loss
    only piecewise smooth, so the comparison is on the gradient as a whole (cosine + norm), not per entry."""
    g = ph.Golden()
    d = g.cam('hyp_cam')
    vol, w = g['o2c.vol'], g['o2c.w'].double()
    base = {k: d[k].double().clone() for k in ('log_quaternion', 'translation', 'viewport')}

    def loss(lq, tr, vp):
        cams, n = cams_struct(d, lq, tr, vp)
        return float((run_o2c(cref, vol, cams, n) * w).sum())

    fd, ref = [], []
    for name, eps in (('log_quaternion', 1e-6), ('translation', 1e-6), ('viewport', 1e-5)):
        for i in range(base[name].shape[0]):
            for j in range(base[name].shape[1]):
                args_p = {k: v.clone() for k, v in base.items()}
                args_m = {k: v.clone() for k, v in base.items()}
                args_p[name][i, j] += eps
                args_m[name][i, j] -= eps
                fd.append((loss(args_p['log_quaternion'], args_p['translation'], args_p['viewport'])
                           - loss(args_m['log_quaternion'], args_m['translation'], args_m['viewport'])) / (2 * eps))
                ref.append(float(g[f'o2c.grad_{name}'][i, j]))
    fd, ref = torch.tensor(fd), torch.tensor(ref)
    cos = torch.dot(fd, ref) / (fd.norm() * ref.norm())
    assert cos > 0.999, (cos, fd, ref)
    assert abs(float(fd.norm() / ref.norm()) - 1.0) < 0.02
