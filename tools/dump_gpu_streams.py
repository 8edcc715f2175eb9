"""Run on a B200: encode small GOPs with the CUDA encoder and save bitstreams + the encoder's reconstructions, so that
the REFERENCE's decoder can be run on them where the reference tree lives (oracle/check_reference_decodes.py).
    python tools/dump_gpu_streams.py gpurun_out"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aivc_b200 import models, gop as G                   # noqa: E402
from aivc_b200.codec import FrameCodec, planes_to_device  # noqa: E402
from aivc_b200.plan import Config                          # noqa: E402
from tests import synth                                    # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out'
dev = torch.device('cuda:0')
CASES = [
    ('bubbles240_gop2', dict(seed=7, C=128, Cy=64, Cz=64, Csc=64, hyper_boost=(12.0, 8.0)), 240, 416, '1_GOP_2', 'bubbles'),
    ('synth_270x480_gop4', dict(seed=1234, C=128, Cy=64, Cz=64, Csc=64, hyper_boost=(12.0, 8.0)), 270, 480, '1_GOP_4', 5),
    ('synth_135x241_ldp3', dict(seed=4, C=64, Cy=32, Cz=32, Csc=32, hyper_boost=(12.0, 8.0)), 135, 241, 'LDP_3', 6),
]
for name, model, h, w, gop_name, src in CASES:
    gop = G.generate_gop_struct(gop_name)
    names = sorted(gop, key=lambda f: int(f.split('_')[1]))
    if src == 'bubbles':              # the real frame, then two shifted copies of it (a moving picture)
        d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden',
                                 'bubbles_416x240_frame0.npz'))
        clip = [(np.roll(d['y'], 2 * t, 1), np.roll(d['u'], t, 1), np.roll(d['v'], t, 1)) for t in range(len(names))]
    else:
        clip = synth.clip(src, len(names), h, w)
    frames = {f: planes_to_device(clip[i], dev) for i, f in enumerate(names)}
    net = models.build_standin(**model)
    for prec in ('bf16x3', 'bf16'):
        codec = FrameCodec(net, h, w, dev, Config(precision=prec))
        bts, rec = codec.encode_gop(frames, gop)
        dec = codec.decode_gop(bts, gop)
        assert all(torch.equal(a, b) for f in names for a, b in zip(rec[f], dec[f]))
        hc, wc = (h + 1) // 2, (w + 1) // 2
        fx = {'H': h, 'W': w, 'gop': gop_name, 'model': repr(model), 'precision': prec}
        for f in names:
            fx['bytes_' + f] = np.frombuffer(bts[f], dtype=np.uint8)
            for k, p, shp in zip('yuv', rec[f], ((h, w), (hc, wc), (hc, wc))):
                fx['rec_%s_%s' % (f, k)] = p.cpu().numpy().reshape(shp)
        np.savez_compressed(os.path.join(out, 'gpu_streams_%s_%s.npz' % (name, prec)), **fx)
        print(name, prec, {f: len(bts[f]) for f in names}, flush=True)
