"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's per-frame quality metrics.

  MSE over Y, U, V            model_mngt/loss_function.py:415-435  (sum of squared errors / number of values)
  PSNR                        loss_function.py:234                 (10 log10(1 / mse), signals in [0, 1])
  MS-SSIM                     func_util/ms_ssim.py:24-150 with val_range = 1 (loss_function.py:443), per
                              plane, weighted by plane size (loss_function.py:454-470)

Pinned by oracle/gen_golden_metrics.py against the reference's own MSELoss / MSSSIMLoss classes imported
from /root/reference/src (tests/golden/metrics.npz).  Only tests/ and bench legs may import this module.
"""
from math import exp

import torch
import torch.nn.functional as F

_W = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)          # ms_ssim.py:98-100


def _window(size):                                      # ms_ssim.py:24-35
    g = torch.Tensor([exp(-(x - size // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(size)])
    g = (g / g.sum()).unsqueeze(1)
    return g.mm(g.t()).float().unsqueeze(0).unsqueeze(0).contiguous()


def _ssim(a, b):                                        # ms_ssim.py:37-92, L = 1, one channel
    win = _window(min(11, a.shape[2], a.shape[3]))
    mu1, mu2 = F.conv2d(a, win), F.conv2d(b, win)
    mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(a * a, win) - mu1_sq
    s2 = F.conv2d(b * b, win) - mu2_sq
    s12 = F.conv2d(a * b, win) - mu12
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    v1, v2 = 2.0 * s12 + c2, s1 + s2 + c2
    return (((2 * mu12 + c1) * v1) / ((mu1_sq + mu2_sq + c1) * v2)).mean(), (v1 / v2).mean()


def msssim_plane(a, b):                                 # ms_ssim.py:95-150
    sims, css = [], []
    for _ in range(5):
        s, c = _ssim(a, b)
        sims.append(s)
        css.append(c)
        pad = (0, a.shape[3] % 2, 0, a.shape[2] % 2)
        a = F.avg_pool2d(F.pad(a, pad, mode='reflect'), (2, 2))
        b = F.avg_pool2d(F.pad(b, pad, mode='reflect'), (2, 2))
    w = torch.FloatTensor(_W)
    pow1, pow2 = torch.stack(css) ** w, torch.stack(sims) ** w
    return torch.prod(pow1[:-1]) * pow2[-1]


def frame_metrics(x_hat, code):
    """x_hat, code: {'y','u','v'} fp32 [1,1,H,W] in [0,1] -> dict(mse, psnr, ms_ssim, ms_ssim_db)."""
    with torch.no_grad():
        n = sum(x_hat[k].numel() for k in 'yuv')
        mse = sum(((x_hat[k] - code[k]) ** 2).sum() for k in 'yuv') / n
        ms = sum(msssim_plane(x_hat[k], code[k]) * x_hat[k].numel() for k in 'yuv') / n
        return {'mse': float(mse), 'psnr': float(10 * torch.log10(1. / mse)), 'ms_ssim': float(ms),
                'ms_ssim_db': float(-10.0 * torch.log10(1 - ms))}
