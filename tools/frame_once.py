"""One I frame + one P frame + one B frame (encode and decode) of the 1080p stand-in: a short
run for ncu launch lists / full captures (bench.py's GOP is 33 frames = ~14k launches)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aivc_b200 import models, gop as G
from aivc_b200.codec import FrameCodec, planes_to_device
from aivc_b200.plan import Config
from bench import synth_gop, MODEL, H, W

dev = torch.device('cuda:0')
net = models.build_standin(**MODEL)
gop = G.generate_gop_struct('1_GOP_2')
codec = FrameCodec(net, H, W, dev, Config(precision=sys.argv[2] if len(sys.argv) > 2 else 'bf16x3'))
clip = synth_gop(3, 3)
frames = {'frame_%d' % i: planes_to_device(clip[i], dev) for i in range(3)}
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(reps):
    bts, rec = codec.encode_gop(frames, gop)
    dec = codec.decode_gop(bts, gop)
torch.cuda.synchronize()
print('ok', {f: len(b) for f, b in bts.items()})
