"""Pins oracle/metrics_ref.py against the reference's own classes (run in the build container, where
/root/reference exists) and writes tests/golden/metrics.npz: seeds + expected values."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/src')
from oracle import metrics_ref as M                       # noqa: E402


def planes(seed, h, w):
    """Two related synthetic frames as uint8 4:2:0 planes (a smooth texture and a noisy copy)."""
    rng = np.random.default_rng(seed)
    out = []
    base = None
    for k, (ph, pw) in enumerate(((h, w), ((h + 1) // 2, (w + 1) // 2), ((h + 1) // 2, (w + 1) // 2))):
        yy, xx = np.mgrid[0:ph, 0:pw]
        a = 127 + 90 * np.sin(xx / (7.0 + k)) * np.cos(yy / (5.0 + 2 * k)) + rng.normal(0, 6, (ph, pw))
        b = a + rng.normal(0, 4 + 2 * k, (ph, pw))
        out.append((np.clip(np.rint(a), 0, 255).astype(np.uint8), np.clip(np.rint(b), 0, 255).astype(np.uint8)))
    return [p[0] for p in out], [p[1] for p in out]


def as_dic(pl):
    return {k: torch.from_numpy(p.astype(np.float32) / 255.)[None, None] for k, p in zip('yuv', pl)}


CASES = [(1, 96, 128), (2, 135, 241), (3, 270, 480)]

if __name__ == '__main__':
    from model_mngt.loss_function import MSELoss, MSSSIMLoss      # the reference's own
    rows = []
    for seed, h, w in CASES:
        a, b = planes(seed, h, w)
        xa, xb = as_dic(a), as_dic(b)
        # get_y_u_v takes the dict form
        mse = float(MSELoss()(xa, xb))
        loss, ms = MSSSIMLoss()(xa, xb)
        ours = M.frame_metrics(xa, xb)
        assert ours['mse'] == mse, (ours['mse'], mse)
        assert ours['ms_ssim'] == float(ms), (ours['ms_ssim'], float(ms))
        rows.append((seed, h, w, mse, float(ms)))
        print(seed, h, w, mse, float(ms), 'oracle == reference')
    np.savez(os.path.join(ROOT, 'tests', 'golden', 'metrics.npz'), cases=np.array(rows, dtype=np.float64))
