"""The reference graph in eager PyTorch ON THE B200 ("reference GPU path"): the stand-in's g_a and
g_s evaluated by the oracle's torch.nn.functional restatement (cuDNN / cuBLAS kernels: the recompiled-library
baseline), next to this repo's fused plans.  Bench-leg use of oracle/."""
import copy, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aivc_b200 import models
from aivc_b200.plan import Plan, Config
from aivc_b200._lib import BF16
from oracle import nn_ref as R
from bench import MODEL, H, W

dev = torch.device('cuda:0')
net = models.build_standin(**MODEL).codec_net.codec_net


def to_dev(m, dtype):
    m = copy.deepcopy(m).to(dev).to(dtype)
    for sub in m.modules():
        if type(sub).__name__ == 'GDN':                 # plain tensor attributes, not buffers (misc_layers.py:78-111)
            for a in ('beta_bound', 'gamma_bound', 'pedestal'):
                setattr(sub, a, getattr(sub, a).to(dev).to(dtype))
    return m


def time_ms(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


cases = {'g_a': (net.g_a, (1, 6, H, W)), 'g_s': (net.g_s, (1, net.nb_ft_y + net.out_c_shortcut_y, 68, 120))}
with torch.no_grad():
    for name, (mod, shape) in cases.items():
        row = []
        for label, dtype, cl in (('fp32 (TF32 convs allowed)', torch.float32, False), ('bf16', torch.bfloat16, False),
                                 ('bf16 channels_last', torch.bfloat16, True)):
            m = to_dev(mod, dtype)
            x = torch.rand(shape, device=dev).to(dtype)
            if cl:
                x = x.contiguous(memory_format=torch.channels_last)
                m = m.to(memory_format=torch.channels_last)
            row.append('%s %.2f ms' % (label, time_ms(lambda: R.forward_module(m, x))))
        plan = Plan(mod, shape[2], shape[3], shape[1], dev, Config(precision='bf16'), out_dtype=BF16, out_pad=0)
        plan.src.buf.t.normal_()
        ours = time_ms(plan.run)
        print('%s at 1080p: eager torch %s | this repo (bf16 plan) %.2f ms = %.0f TFLOP/s'
              % (name, ', '.join(row), ours, plan.flops() / ours / 1e9), flush=True)
