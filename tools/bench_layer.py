"""Micro-benchmark of single fused stages (CUDA events around `reps` back-to-back runs; --flush --reps 1
for cold-L2 single launches).  AIVC_B200_LIB=<other build> gives an A/B on the same box."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aivc_b200.layers as M
from aivc_b200.plan import Plan, Config
from aivc_b200._lib import BF16, BF16X2

class _GdnRes(torch.nn.Module):
    """x + igdn(conv3(x)): the residual flavour of the fused conv + GDN stage (lowered like a ChengResBlock tail)."""
    def __init__(self, inverse):
        super().__init__()
        self.mode = 'plain'
        self.layers = torch.nn.Sequential(M.CustomConvLayer(3, 128, 128, non_linearity='gdn_inverse' if inverse else 'gdn'))
_GdnRes.__name__ = 'ChengResBlock'


CASES = {
    'c3igdn_res_544': (lambda: _GdnRes(True), 128, 544, 960),
    'c3gdn_res_270': (lambda: _GdnRes(False), 128, 270, 480),
    'c3_128_540': (lambda: M.CustomConvLayer(3, 128, 128, non_linearity='leaky_relu'), 128, 540, 960),
    'c3_128_270': (lambda: M.CustomConvLayer(3, 128, 128, non_linearity='leaky_relu'), 128, 270, 480),
    'res_128_270': (lambda: M.ResBlock(3, 128), 128, 270, 480),
    'c3_128_135': (lambda: M.CustomConvLayer(3, 128, 128, non_linearity='leaky_relu'), 128, 135, 240),
    'c3_128_68': (lambda: M.CustomConvLayer(3, 128, 128, non_linearity='leaky_relu'), 128, 68, 120),
    'c3s2_128_540': (lambda: M.CustomConvLayer(3, 128, 128, non_linearity='leaky_relu', conv_stride=2), 128, 540, 960),
    'c3gdn_128_270': (lambda: M.CustomConvLayer(3, 128, 128, non_linearity='gdn'), 128, 270, 480),
    'c3igdn_128_544': (lambda: M.CustomConvLayer(3, 128, 128, non_linearity='gdn_inverse'), 128, 544, 960),
    'up3_128_270': (lambda: M.UpscalingLayer(3, 128, 128, non_linearity='leaky_relu'), 128, 270, 480),
    'up5_128_16_544': (lambda: M.UpscalingLayer(5, 128, 16, non_linearity='no'), 128, 544, 960),
    'cheng_plain_270': (lambda: M.ChengResBlock(128, 'plain'), 128, 270, 480),
    'cheng_down_540': (lambda: M.ChengResBlock(128, 'down'), 128, 540, 960),
    'cheng_up_272': (lambda: M.ChengResBlock(128, 'up_tconv'), 128, 272, 480),
    'att_128_270': (lambda: M.SimplifiedAttention(128), 128, 270, 480),
    'att_64_68': (lambda: M.SimplifiedAttention(64), 64, 68, 120),
    'att_128_68': (lambda: M.SimplifiedAttention(128), 128, 68, 120),
    'up5_128_6_544': (lambda: M.UpscalingLayer(5, 128, 6, non_linearity='no'), 128, 544, 960),
    'up5_128_3_544': (lambda: M.UpscalingLayer(5, 128, 3, non_linearity='no'), 128, 544, 960),
    'c5s2_16_128_1080': (lambda: M.CustomConvLayer(5, 16, 128, non_linearity='gdn', conv_stride=2), 16, 1080, 1920),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cases', default=','.join(CASES))
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--reps', type=int, default=20, help='back-to-back runs per timing sample (L2-warm, as inside a transform)')
    ap.add_argument('--flush', action='store_true', help='flush L2 before every sample (use with --reps 1)')
    ap.add_argument('--precision', default='bf16x3', choices=['bf16x3', 'bf16'])
    ap.add_argument('--f32-out', action='store_true', help='fp32 un-bordered output (default: bf16 + 1-pixel border, as inside a transform)')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name in args.cases.split(','):
        mk, cin, h, w = CASES[name]
        torch.manual_seed(0)
        act = BF16X2 if args.precision == 'bf16x3' else BF16
        plan = Plan(mk().eval(), h, w, cin, dev, Config(precision=args.precision), out_dtype=None if args.f32_out else act, out_pad=1)
        plan.src.buf.t.normal_()
        for _ in range(3):
            plan.run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            if args.flush:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.reps):
                plan.run()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / args.reps)
        ts.sort()
        med = ts[len(ts) // 2]
        print('%-18s stages=%d  median %8.1f us  min %8.1f us  %7.1f TFLOP/s (algorithmic)'
              % (name, len(plan.stages), med * 1e3, ts[0] * 1e3, plan.flops() / (med * 1e-3) / 1e12), flush=True)


if __name__ == '__main__':
    main()
