"""Micro-benchmark of single fused stages (CUDA events, L2 flushed between launches)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aivc_b200.layers as M
from aivc_b200.plan import Plan, Config

CASES = {
    'c3_128_540': (lambda: M.CustomConvLayer(3, 128, 128, non_linearity='leaky_relu'), 128, 540, 960),
    'c3_128_270': (lambda: M.CustomConvLayer(3, 128, 128, non_linearity='leaky_relu'), 128, 270, 480),
    'res_128_270': (lambda: M.ResBlock(3, 128), 128, 270, 480),
    'c3_128_135': (lambda: M.CustomConvLayer(3, 128, 128, non_linearity='leaky_relu'), 128, 135, 240),
    'c3_128_68': (lambda: M.CustomConvLayer(3, 128, 128, non_linearity='leaky_relu'), 128, 68, 120),
    'c3s2_128_540': (lambda: M.CustomConvLayer(3, 128, 128, non_linearity='leaky_relu', conv_stride=2), 128, 540, 960),
    'c3gdn_128_270': (lambda: M.CustomConvLayer(3, 128, 128, non_linearity='gdn'), 128, 270, 480),
    'up3_128_270': (lambda: M.UpscalingLayer(3, 128, 128, non_linearity='leaky_relu'), 128, 270, 480),
    'up5_128_16_544': (lambda: M.UpscalingLayer(5, 128, 16, non_linearity='no'), 128, 544, 960),
    'c5s2_16_128_1080': (lambda: M.CustomConvLayer(5, 16, 128, non_linearity='gdn', conv_stride=2), 16, 1080, 1920),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cases', default=','.join(CASES))
    ap.add_argument('--iters', type=int, default=10)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name in args.cases.split(','):
        mk, cin, h, w = CASES[name]
        torch.manual_seed(0)
        plan = Plan(mk().eval(), h, w, cin, dev, Config(precision='bf16'))
        plan.src.buf.t.normal_()
        for _ in range(3):
            plan.run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            plan.run()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        med = ts[len(ts) // 2]
        print('%-18s stages=%d  median %8.1f us  min %8.1f us  %7.1f TFLOP/s (algorithmic)'
              % (name, len(plan.stages), med * 1e3, ts[0] * 1e3, plan.flops() / (med * 1e-3) / 1e12), flush=True)


if __name__ == '__main__':
    main()
