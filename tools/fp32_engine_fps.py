"""One-off: the exact fp32 (SIMT) engine at the benchmark's resolution.  One frame of every type (1080p I, P, B) is
encoded and decoded, the 33-frame GOP figure is composed by frame-type counts like the CPU arm's.  Prints JSON."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aivc_b200 import models, gop as G
from aivc_b200.codec import FrameCodec
from aivc_b200.plan import Config
from bench import synth_gop, MODEL, H, W

dev = torch.device('cuda:0')
net = models.build_standin(**MODEL)
codec = FrameCodec(net, H, W, dev, Config(precision='fp32'))
clip = synth_gop(100, 3)
planes = [tuple(torch.from_numpy(p.reshape(-1)).to(dev) for p in fr) for fr in clip]
t = {}
rec = {}
for rep in range(2):                               # first pass = warm-up
    for name, ft, idx, refs in (('I', 0, 0, (None, None)), ('P', 1, 2, (0, None)), ('B', 2, 1, (0, 2))):
        prev = rec.get(refs[0]) if refs[0] is not None else None
        nxt = rec.get(refs[1]) if refs[1] is not None else None
        torch.cuda.synchronize(); t0 = time.perf_counter()
        data, r = codec.encode_frame(planes[idx], ft, prev, nxt)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        d = codec.decode_frame(data, ft, prev, nxt)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        assert all(torch.equal(a, b) for a, b in zip(r, d))
        rec[idx] = r
        t[name] = (t1 - t0, t2 - t1)
gop_s = sum(t['I']) + sum(t['P']) + 31 * sum(t['B'])
print(json.dumps({'engine': 'fp32 (exact SIMT)', 'seconds_enc_dec': {k: [round(x, 3) for x in v] for k, v in t.items()},
                  'gop_seconds': round(gop_s, 2), 'frames_per_s': round(33 / gop_s, 3),
                  'how': '1080p I, P, B frame encode+decode (frame-serial API), GOP of 33 = t_I + t_P + 31 t_B'}))
