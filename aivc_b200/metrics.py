"""Per-frame quality metrics on the device: what the reference's encode loop logs for
every frame through compute_metrics_one_GOP (model_mngt/loss_function.py:103-257) -- MSE / PSNR over the
Y, U, V planes and the plane-size-weighted MS-SSIM -- computed from uint8 4:2:0 planes by
`aivc_frame_metrics` (csrc/metrics.cu).  No CPU fallback."""
import ctypes as C

import torch

from . import _lib

_SCRATCH = {}


def frame_metrics_async(planes_a, planes_b, h, w):
    """planes_*: (y, u, v) flat uint8 CUDA tensors of one h x w 4:2:0 frame.  Enqueues the kernels on the
    current stream and returns a 4-element fp32 CUDA tensor: mse, psnr, ms_ssim, ms_ssim_db."""
    dev = planes_a[0].device
    if dev.type != 'cuda':
        raise RuntimeError('aivc_b200.metrics runs on a CUDA device only (no CPU fallback)')
    for p in tuple(planes_a) + tuple(planes_b):
        if p.dtype != torch.uint8 or not p.is_contiguous() or p.device != dev:
            raise ValueError('planes must be contiguous uint8 tensors on one device')
    hc, wc = (h + 1) // 2, (w + 1) // 2
    if planes_a[0].numel() != h * w or planes_a[1].numel() != hc * wc or planes_b[0].numel() != h * w:
        raise ValueError('plane sizes do not match %dx%d 4:2:0' % (w, h))
    L = _lib.lib()
    key = (dev.index, h, w)
    sc = _SCRATCH.get(key)
    if sc is None:
        sc = _SCRATCH[key] = torch.empty(L.aivc_frame_metrics_scratch_bytes(h, w), dtype=torch.uint8, device=dev)
    out = torch.empty(4, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.aivc_frame_metrics(*(p.data_ptr() for p in planes_a), *(p.data_ptr() for p in planes_b), h, w,
                                        sc.data_ptr(), sc.numel(), out.data_ptr(), _lib.stream_ptr()))
    return out


def frame_metrics(planes_a, planes_b, h, w):
    """-> {'mse', 'psnr', 'ms_ssim', 'ms_ssim_db'} (python floats; synchronises)."""
    v = frame_metrics_async(planes_a, planes_b, h, w).cpu().tolist()
    return dict(zip(('mse', 'psnr', 'ms_ssim', 'ms_ssim_db'), v))
