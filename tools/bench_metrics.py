"""Device time of the per-frame metrics (MSE / PSNR / MS-SSIM) at 1080p, CUDA events, against the oracle on
the host cores (bench-leg use of oracle/)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aivc_b200 import metrics
from oracle import metrics_ref as M, gen_golden_metrics as Gm

h, w = 1080, 1920
a, b = Gm.planes(7, h, w)
dev = torch.device('cuda:0')
to_dev = lambda pl: tuple(torch.from_numpy(np.ascontiguousarray(p).reshape(-1)).to(dev) for p in pl)
da, db = to_dev(a), to_dev(b)
for _ in range(3):
    metrics.frame_metrics_async(da, db, h, w)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = metrics.frame_metrics_async(da, db, h, w)
e1.record()
torch.cuda.synchronize()
got = dict(zip(('mse', 'psnr', 'ms_ssim', 'ms_ssim_db'), out.cpu().tolist()))
torch.set_num_threads(os.cpu_count())
t0 = time.time()
ref = M.frame_metrics(Gm.as_dic(a), Gm.as_dic(b))
t_cpu = time.time() - t0
print('frame_metrics 1080p: %.1f us per frame on the GPU; oracle %.0f ms on %d host cores; ms_ssim %.7f vs %.7f, psnr %.4f vs %.4f'
      % (e0.elapsed_time(e1) * 1e3 / 20, t_cpu * 1e3, os.cpu_count(), got['ms_ssim'], ref['ms_ssim'], got['psnr'], ref['psnr']))
