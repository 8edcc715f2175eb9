"""Parity of the CUDA path against the CPU oracle beyond the golden fixtures (bench-leg use of oracle/):
BASELINE config 1 geometry (416x240, all-intra frame 0 + the I/P/B GOP '1_GOP_2'), stand-in model C=128,
synthetic frames.  Reports, for the fp32 (exact) and bf16 (tcgen05) engines: quantised-index mismatches
of the I frame's latents, bitstream bytes per frame, PSNR of our reconstruction against the oracle's and
the PSNR-vs-source delta.  Prints one JSON line."""
import json, os, sys, time
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aivc_b200 import models, gop as G
from aivc_b200.codec import FrameCodec, planes_to_device
from aivc_b200.plan import Config
from bench import synth_gop, MODEL
from oracle import codec_ref as O

H, W = 240, 416
dev = torch.device('cuda:0')
net = models.build_standin(**MODEL)
tables = O.Tables(net)
gop = G.generate_gop_struct('1_GOP_2')
clip = synth_gop(5, 3, H, W)
names = ['frame_0', 'frame_1', 'frame_2']


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def planes_of(rec_dic):
    return [np.rint(rec_dic[k].numpy() * 255).astype(np.uint8).reshape(-1) for k in 'yuv']


torch.set_num_threads(os.cpu_count())
t0 = time.time()
yuv = {f: {k: torch.from_numpy(p.astype(np.float32) / 255.)[None, None] for k, p in zip('yuv', clip[i])}
       for i, f in enumerate(names)}
# oracle: I frame with aux (latents), then the whole GOP
z = O.zero_yuv(H, W)
_, _, aux_i = O.encode_frame(net, tables, yuv['frame_0'], z, z, 0)
o_bytes, o_rec = O.encode_gop(net, tables, yuv, gop)
t_oracle = time.time() - t0
src = {f: [p.reshape(-1) for p in clip[i]] for i, f in enumerate(names)}
out = {'geometry': '%dx%d, GOP 1_GOP_2 (I, P, B), stand-in C=128 Cy=Cz=64, synthetic frames' % (W, H),
       'oracle_seconds': round(t_oracle, 1), 'oracle_bytes': {f: len(o_bytes[f]) for f in names}}
frames = {f: planes_to_device(clip[i], dev) for i, f in enumerate(names)}
for prec in ('fp32', 'bf16x3', 'bf16'):
    codec = FrameCodec(net, H, W, dev, Config(precision=prec))
    # I frame alone: latent indices of the CodecNet
    codec.encode_gop({'frame_0': frames['frame_0']}, G.generate_gop_struct('1_GOP_0'))
    torch.cuda.synchronize()
    q = codec.codec.q_dev.cpu().numpy().astype(np.int32)
    zq = codec.codec.z_dev.cpu().numpy().astype(np.int32)
    q_ref = aux_i['codec']['q'].numpy().astype(np.int32).reshape(-1)
    z_ref = aux_i['codec']['z_hat'].numpy().astype(np.int32).reshape(-1)
    bts, rec = codec.encode_gop(frames, gop)
    dec = codec.decode_gop(bts, gop)
    closed = all(torch.equal(a, b) for f in names for a, b in zip(rec[f], dec[f]))
    r = {'y_index_mismatches': int((q != q_ref).sum()), 'y_symbols': int(q.size),
         'z_index_mismatches': int((zq != z_ref).sum()), 'z_symbols': int(zq.size),
         'closed_loop_exact': bool(closed), 'bytes': {f: len(bts[f]) for f in names},
         'bytes_identical': {f: bts[f] == o_bytes[f] for f in names}, 'frames': {}}
    for f in names:
        ours = [p.cpu().numpy() for p in rec[f]]
        orc = planes_of(o_rec[f])
        cat = lambda ps: np.concatenate(ps)
        r['frames'][f] = {'psnr_ours_vs_oracle_db': round(psnr(cat(ours), cat(orc)), 3),
                          'max_abs_level_diff': int(np.abs(cat(ours).astype(int) - cat(orc).astype(int)).max()),
                          'psnr_vs_source_ours_db': round(psnr(cat(ours), cat(src[f])), 4),
                          'psnr_vs_source_oracle_db': round(psnr(cat(orc), cat(src[f])), 4)}
    out[prec] = r
print(json.dumps(out))
