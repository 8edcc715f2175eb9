"""Is the launch path host-bound?  For every transform plan of the 1080p stand-in: wall time the
host needs to ENQUEUE the plan (no sync) against the time the GPU needs to run it (CUDA events)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aivc_b200 import models
from aivc_b200.codec import FrameCodec
from aivc_b200.plan import Config
from bench import MODEL, H, W

dev = torch.device('cuda:0')
net = models.build_standin(**MODEL)
codec = FrameCodec(net, H, W, dev, Config(precision='bf16'))
reps = 4
for eng_name in ('mof', 'codec'):
    eng = getattr(codec, eng_name)
    for pn in ('g_a', 'g_a_ref', 'h_a', 'h_s', 'g_s'):
        plan = getattr(eng, pn, None)
        if plan is None:
            continue
        for _ in range(2):
            plan.run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for _ in range(reps):
            plan.run()
        b.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        n = len(plan.stages)
        host = (t1 - t0) * 1e6 / reps
        gpu = a.elapsed_time(b) * 1e3 / reps
        print('%-5s %-8s stages %3d  host enqueue %8.1f us (%5.2f us/stage)   gpu %8.1f us (%6.2f us/stage)  %6.1f TFLOP/s'
              % (eng_name, pn, n, host, host / n, gpu, gpu / n, plan.flops() / gpu / 1e6), flush=True)
