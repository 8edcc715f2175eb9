"""Where does decode_gop spend its wall time?  (host phases vs GPU)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aivc_b200 import models, gop as G
from aivc_b200.codec import FrameCodec, planes_to_device
from aivc_b200.plan import Config
from bench import synth_gop, MODEL, H, W

dev = torch.device('cuda:0')
net = models.build_standin(**MODEL)
gop = G.generate_gop_struct('1_GOP_32')
names = sorted(gop, key=lambda f: int(f.split('_')[1]))
codec = FrameCodec(net, H, W, dev, Config(precision='bf16'))
clip = synth_gop(100, len(names))
frames = {f: planes_to_device(clip[i], dev) for i, f in enumerate(names)}
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    bts, rec = codec.encode_gop(frames, gop)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    dec = codec.decode_gop(bts, gop)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print('rep %d: encode %.1f ms, decode %.1f ms' % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1)), flush=True)

