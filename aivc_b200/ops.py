"""Tensor-level entry points behind the drop-in modules (NCHW fp32 CUDA tensors in/out).

Each function is a thin marshalling layer over one C-ABI call; the arithmetic is in
aivc_b200/csrc/pointwise.cu.  CUDA tensors only -- no CPU fallback.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import FMap, F32


def _need_cuda(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError('aivc_b200.ops run on CUDA tensors only (no CPU fallback)')


def nchw_fmap(t):
    """FMap describing a contiguous [1,C,H,W]... no: NHWC [H,W,C] fp32 tensor without border."""
    h, w, c = t.shape
    return FMap(t.data_ptr(), h, w, c, 0, c, 0, w, h, F32, 0)


def mu_sigma(x, nb_ft):
    """PdfParamParameterizer (misc_layers.py:180-269), single-component mode."""
    _need_cuda(x)
    x = x.contiguous().float()
    b, c2, h, w = x.shape
    if b != 1 or c2 < 2 * nb_ft:
        raise ValueError('mu_sigma: expected [1, >=2*nb_ft, H, W]')
    mu = torch.empty((1, nb_ft, h, w), device=x.device, dtype=torch.float32)
    sigma = torch.empty_like(mu)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().aivc_mu_sigma_nchw(x.data_ptr(), mu.data_ptr(), sigma.data_ptr(), nb_ft,
                                                 h * w, _lib.stream_ptr()))
    return mu, sigma


def pdf_prob(y, mu, sigma, family, out=None):
    """One component of ParametricPdf.forward (pdf_estimator.py:27-70): cdf(y + 1/2) - cdf(y - 1/2);
    `out` given: accumulate into it."""
    _need_cuda(y, sigma)
    y, sigma = y.contiguous().float(), sigma.contiguous().float()
    mu = None if mu is None else mu.contiguous().float()
    acc = out is not None
    if out is None:
        out = torch.empty_like(y)
    with torch.cuda.device(y.device):
        _lib.check(_lib.lib().aivc_pdf_prob(y.data_ptr(), None if mu is None else mu.data_ptr(), sigma.data_ptr(),
                                            {'laplace': 0, 'normal': 1}[family], 1 if acc else 0, out.data_ptr(),
                                            y.numel(), _lib.stream_ptr()))
    return out


def yuv420_to_444(y, u, v):
    """InputLayer.forward (ae_layers.py:27-35) -> [1,3,H,W]."""
    _need_cuda(y, u, v)
    h, w = y.shape[2:]
    L = _lib.lib()
    with torch.cuda.device(y.device):
        st = _lib.stream_ptr()
        nhwc = torch.empty((h, w, 3), device=y.device, dtype=torch.float32)
        fm = nchw_fmap(nhwc)
        _lib.check(L.aivc_yuv420_to_fmap(y.contiguous().float().data_ptr(), u.contiguous().float().data_ptr(),
                                         v.contiguous().float().data_ptr(), 0, 0, C.byref(fm), st))
        out = torch.empty((1, 3, h, w), device=y.device, dtype=torch.float32)
        _lib.check(L.aivc_fmap_to_nchw(C.byref(fm), out.data_ptr(), st))
    return out


def yuv444_to_420(x):
    """OutputLayer.forward (ae_layers.py:42-56): Y = ch 0, U,V = bilinear x0.5 of ch 1,2
    (floor(H/2) x floor(W/2), as torch's interpolate returns)."""
    _need_cuda(x)
    # the unquantised 2x2 mean is also what aivc_finalize_frame computes before its 8-bit cast;
    # this module-level variant keeps fp32 like the reference class does.
    _, _, h, w = x.shape
    x = x.contiguous().float()
    a = x[:, 1:, 0:2 * (h // 2):2, 0:2 * (w // 2):2]
    b = x[:, 1:, 0:2 * (h // 2):2, 1:2 * (w // 2):2]
    c = x[:, 1:, 1:2 * (h // 2):2, 0:2 * (w // 2):2]
    d = x[:, 1:, 1:2 * (h // 2):2, 1:2 * (w // 2):2]
    uv = 0.5 * (0.5 * a + 0.5 * b) + 0.5 * (0.5 * c + 0.5 * d)
    return x[:, 0:1], uv[:, 0:1], uv[:, 1:2]


def channel_scale(x, g):
    """GainMatrix.forward: x * |gain| (gain_matrix.py:122-124); per-channel broadcast."""
    _need_cuda(x)
    return x * g.to(x.device).view(1, -1, 1, 1)


def warp_blend(prev, nxt, v_prev, v_next, beta):
    """motion_compensation contract (decode.py:524-533) on NCHW fp32 tensors."""
    _need_cuda(prev, nxt, v_prev, v_next, beta)
    _, _, h, w = prev.shape
    args = [t.contiguous().float() for t in (prev, nxt, v_prev, v_next, beta)]
    out = torch.empty((1, 3, h, w), device=prev.device, dtype=torch.float32)
    with torch.cuda.device(prev.device):
        _lib.check(_lib.lib().aivc_warp_blend_nchw(*[a.data_ptr() for a in args], out.data_ptr(), h, w,
                                                   _lib.stream_ptr()))
    return out
