"""A/B: frames in flight (1 | 2), both tensor-core precisions.  One 1080p RA GOP encode + decode, resident inputs.
Measured on B200 (round 2): bf16x3 23.6 -> 23.9 frames/s, bf16 60.5 -> 61.3 (with programmatic dependent launch
switched off as well: 23.8 / 61.6); after the epilogue rework: bf16x3 28.1 -> 28.4 -- the device is 95 % busy and
power-capped with one frame in flight."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aivc_b200 import models, gop as G, _lib
from aivc_b200.codec import FrameCodec
from aivc_b200.plan import Config
from bench import synth_gop, MODEL, H, W, GOP_NAME

dev = torch.device('cuda:0')
net = models.build_standin(**MODEL)
gop = G.generate_gop_struct(GOP_NAME)
names = sorted(gop, key=lambda f: int(f.split('_')[1]))
clip = synth_gop(100, len(names))
frames = {f: tuple(torch.from_numpy(p.reshape(-1)).to(dev) for p in clip[i]) for i, f in enumerate(names)}
L = _lib.lib()
for prec in sys.argv[1:] or ['bf16x3', 'bf16']:
    ref = None
    for fif in (1, 2):
        codec = FrameCodec(net, H, W, dev, Config(precision=prec, frames_in_flight=fif))
        for pdl in (1,):
            for _ in range(2):
                bts, rec = codec.encode_gop(frames, gop)
                dec = codec.decode_gop(bts, gop)
            torch.cuda.synchronize()
            te = td = 0.0
            for _ in range(2):
                t0 = time.perf_counter(); bts, rec = codec.encode_gop(frames, gop); torch.cuda.synchronize()
                t1 = time.perf_counter(); dec = codec.decode_gop(bts, gop); torch.cuda.synchronize()
                te += t1 - t0; td += time.perf_counter() - t1
            ok = all(torch.equal(a, b) for f in names for a, b in zip(rec[f], dec[f]))
            if ref is None:
                ref = bts
            print('%s frames_in_flight %d pdl %d: enc %.0f ms dec %.0f ms -> %.1f fps, closed loop %s, bytes identical to first config %s'
                  % (prec, fif, pdl, 500 * te, 500 * td, 66 / (te + td), ok, bts == ref), flush=True)
        del codec
        torch.cuda.empty_cache()
