"""Development check of the bf16x3 (split-bf16 tcgen05) mode: error statistics of leaves / wide layers against the
oracle, and index / byte parity of the 80x112 golden GOP.  Run on a B200: python tools/x3_check.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.leafcfg import LEAVES, load_leaf           # noqa: E402
from tests.test_gpu_engine_tc import WIDE              # noqa: E402


def stats(y, ref):
    err = y - ref
    rms = np.sqrt((err ** 2).mean()) / max(np.sqrt((ref ** 2).mean()), 1e-12)
    mx = np.abs(err).max() / max(np.abs(ref).max(), 1e-12)
    return rms, mx


def main():
    from aivc_b200 import plan, models, gop as G
    import aivc_b200.layers as M
    from aivc_b200.plan import Config
    from aivc_b200.codec import FrameCodec, planes_to_device
    from oracle import nn_ref as R
    dev = torch.device('cuda:0')
    gd = os.path.join(ROOT, 'tests', 'golden')
    for prec in ('bf16x3', 'fp32'):
        cfg = Config(precision=prec)
        print('==== precision', prec)
        for name in sorted(LEAVES):
            m, fx = load_leaf(name, gd)
            y = plan.run_module(m, torch.from_numpy(fx['x0']).to(dev), cfg).cpu().numpy()
            print('leaf %-18s rms %.2e max %.2e' % ((name,) + stats(y, fx['y0'])))
        for name in sorted(WIDE):
            mk, cin, _ = WIDE[name]
            torch.manual_seed(hash(name) % 1000)
            m = mk(M).eval()
            for h, w in ((33, 47), (135, 243)):
                x = torch.randn(1, cin, h, w, generator=torch.Generator().manual_seed(5))
                with torch.no_grad():
                    ref = R.forward_module(m, x.double() if False else x).numpy()
                y = plan.run_module(m, x.to(dev), cfg).cpu().numpy()
                p = next(iter(plan.cached_plans(m).values()))[1]
                eng = [s.engine for s in p.stages]
                plan.cached_plans(m).clear()
                print('wide %-20s %3dx%-3d rms %.2e max %.2e engines %s' % ((name, h, w) + stats(y, ref) + (eng,)))
        fx = np.load(os.path.join(gd, 'system_80x112.npz'))
        h, w = int(fx['H']), int(fx['W'])
        net = models.build_standin(seed=4321, C=32, Cy=16, Cz=16, Csc=16)
        gop = G.generate_gop_struct('1_GOP_2')
        frames = {'frame_%d' % t: planes_to_device([fx['src_frame_%d_%s' % (t, k)] for k in 'yuv'], dev) for t in range(3)}
        codec = FrameCodec(net, h, w, dev, cfg)
        bts, rec = codec.encode_gop(frames, gop)
        dec = codec.decode_gop(bts, gop)
        for f in sorted(gop):
            same = bts[f] == fx['spec_bytes_%s' % f].tobytes()
            lev = max(int(np.abs(p.cpu().numpy().astype(np.int32) - fx['spec_rec_%s_%s' % (f, k)].reshape(-1).astype(np.int32)).max())
                      for k, p in zip('yuv', rec[f]))
            closed = all(torch.equal(a, b) for a, b in zip(rec[f], dec[f]))
            print('system %s: bytes identical %s (%d vs %d), max level diff %d, closed loop %s'
                  % (f, same, len(bts[f]), len(fx['spec_bytes_%s' % f].tobytes()), lev, closed))


if __name__ == '__main__':
    main()
