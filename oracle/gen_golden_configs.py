"""ORACLE (test infrastructure): golden fixtures at the BASELINE.json configurations, minted by the CPU oracle
(oracle/codec_ref.py, pinned to the reference classes by oracle/gen_golden.py) in this container.

    python -m oracle.gen_golden_configs [case ...]        # writes tests/golden/cfg_<case>.npz

  bubbles240 : configs[0] -- frame 0 of raw_videos/BlowingBubbles_416x240_50_420.yuv (src/sanity_script.sh:1-13),
               all intra ('1_GOP_0'), stand-in seed 7 ("ms_ssim-7"), C=128
  ldp720     : configs[1] -- synthetic 1280x720, low-delay P, frames I, P, P ('LDP_2' = the first three frames of
               an LDP_8 GOP in coding order), stand-in seed 4, C=128
  ra1080     : configs[2] -- synthetic 1920x1080, random access I, P, B ('1_GOP_2': every frame type of 1_GOP_32),
               stand-in seed 1234 (the benchmark model), C=128
  ra1080_gop8: the same model on a 9-frame random-access GOP ('1_GOP_8': B frames three levels deep)

Per frame the fixture keeps: the quantised latent indices of both nets (int8/int16, what north_star asks to be
bit-exact), the z indices, length + md5 of the frame's bitstream bytes, md5 of the reconstructed planes, a 1/16
subsample of them (every 4th row and column: bounds level differences without 3 MB per frame) and the PSNR against
the source.  Sources: tests/synth.py (integer-exact) or the committed first frame of the real clip."""
import hashlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aivc_b200 import models, gop as G          # noqa: E402  (layer mirrors: parameters only; the oracle evaluates them)
from oracle import codec_ref as O               # noqa: E402
from tests import synth                         # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
CASES = {
    'bubbles240': dict(h=240, w=416, gop='1_GOP_0', model=dict(seed=7, C=128, Cy=64, Cz=64, Csc=64, hyper_boost=(12.0, 8.0)), src='bubbles'),
    'ldp720': dict(h=720, w=1280, gop='LDP_2', model=dict(seed=4, C=128, Cy=64, Cz=64, Csc=64, hyper_boost=(12.0, 8.0)), src=('synth', 720)),
    'ra1080': dict(h=1080, w=1920, gop='1_GOP_2', model=dict(seed=1234, C=128, Cy=64, Cz=64, Csc=64, hyper_boost=(12.0, 8.0)), src=('synth', 1080)),
    # a deeper random-access hierarchy (9 frames, dependency levels of 1, 1, 1, 2, 4 frames): does the distance to the
    # oracle grow along the reference chain?
    'ra1080_gop8': dict(h=1080, w=1920, gop='1_GOP_8', model=dict(seed=1234, C=128, Cy=64, Cz=64, Csc=64, hyper_boost=(12.0, 8.0)), src=('synth', 1081)),
}


def source_frames(case, n):
    c = CASES[case]
    h, w = c['h'], c['w']
    if c['src'] == 'bubbles':
        p = os.path.join(OUT, 'bubbles_416x240_frame0.npz')
        if not os.path.exists(p):          # one-off: first frame of the reference's own test clip
            raw = np.fromfile('/root/reference/raw_videos/BlowingBubbles_416x240_50_420.yuv', dtype=np.uint8,
                              count=w * h * 3 // 2)
            np.savez_compressed(p, y=raw[:w * h].reshape(h, w), u=raw[w * h:w * h * 5 // 4].reshape(h // 2, w // 2),
                                v=raw[w * h * 5 // 4:].reshape(h // 2, w // 2))
        d = np.load(p)
        return [(d['y'], d['u'], d['v'])]
    return synth.clip(c['src'][1], n, h, w)


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else float(10 * np.log10(255.0 ** 2 / mse))


def compact(q):
    q = q.numpy().astype(np.int32).reshape(q.shape[1:])
    return q.astype(np.int8) if -128 <= q.min() and q.max() <= 127 else q.astype(np.int16)


def mint(case):
    c = CASES[case]
    h, w = c['h'], c['w']
    torch.set_num_threads(os.cpu_count())
    net = models.build_standin(**c['model'])
    tables = O.Tables(net)
    gop = G.generate_gop_struct(c['gop'])
    names = sorted(gop, key=lambda f: int(f.split('_')[1]))
    clip = source_frames(case, len(names))
    yuv = {f: {k: torch.from_numpy(p.astype(np.float32) / 255.)[None, None] for k, p in zip('yuv', clip[i])}
           for i, f in enumerate(names)}
    fx = {'H': h, 'W': w, 'gop': c['gop'], 'model': repr(sorted(c['model'].items()))}
    rec = {}
    for f in sorted(gop, key=lambda f: gop[f]['coding_order']):
        t0 = time.time()
        t = gop[f]['type']
        prev = rec[gop[f]['prev_ref']] if t != 0 else O.zero_yuv(h, w)
        nxt = rec[gop[f]['next_ref']] if t == 2 else O.zero_yuv(h, w)
        data, rec[f], aux = O.encode_frame(net, tables, yuv[f], prev, nxt, t)
        fx[f + '_type'] = t
        fx[f + '_nbytes'] = len(data)
        fx[f + '_bytes_md5'] = hashlib.md5(data).hexdigest()
        fx[f + '_sections'] = np.array([len(s) for s in O.split_frame_sections(data)])
        for net_name in ('mof', 'codec'):
            if net_name in aux:
                fx['%s_%s_q' % (f, net_name)] = compact(aux[net_name]['q'])
                fx['%s_%s_z' % (f, net_name)] = compact(aux[net_name]['z_hat'])
        planes = [np.rint(rec[f][k].numpy() * 255).astype(np.uint8)[0, 0] for k in 'yuv']
        src = clip[names.index(f)]
        fx[f + '_planes_md5'] = hashlib.md5(b''.join(p.tobytes() for p in planes)).hexdigest()
        fx[f + '_psnr_vs_source'] = psnr(np.concatenate([p.reshape(-1) for p in planes]),
                                         np.concatenate([p.reshape(-1) for p in src]))
        for k, p in zip('yuv', planes):
            fx['%s_sub_%s' % (f, k)] = p[::4, ::4].copy()
        print('%s %s type %d: %d bytes, psnr %.4f dB, %.1f s' % (case, f, t, len(data), fx[f + '_psnr_vs_source'],
                                                                 time.time() - t0), flush=True)
    np.savez_compressed(os.path.join(OUT, 'cfg_%s.npz' % case), **fx)


if __name__ == '__main__':
    for case in (sys.argv[1:] or list(CASES)):
        mint(case)
