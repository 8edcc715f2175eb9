"""Shared by tests/test_gpu_parity_configs.py, tools/parity_configs.py and bench.py's parity leg: run the CUDA codec
on one of the BASELINE-configuration fixtures (oracle/gen_golden_configs.py) and compare with what the CPU oracle
produced: quantised latent indices, z indices, bitstream bytes, reconstructed planes, PSNR against the source."""
import hashlib
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
MODELS = {
    'bubbles240': dict(seed=7, C=128, Cy=64, Cz=64, Csc=64, hyper_boost=(12.0, 8.0)),
    'ldp720': dict(seed=4, C=128, Cy=64, Cz=64, Csc=64, hyper_boost=(12.0, 8.0)),
    'ra1080': dict(seed=1234, C=128, Cy=64, Cz=64, Csc=64, hyper_boost=(12.0, 8.0)),
    'ra1080_gop8': dict(seed=1234, C=128, Cy=64, Cz=64, Csc=64, hyper_boost=(12.0, 8.0)),
}
SYNTH_SEED = {'ldp720': 720, 'ra1080': 1080, 'ra1080_gop8': 1081}


def source_frames(case, fx):
    from tests import synth
    h, w = int(fx['H']), int(fx['W'])
    if case == 'bubbles240':
        d = np.load(os.path.join(GOLDEN, 'bubbles_416x240_frame0.npz'))
        return [(d['y'], d['u'], d['v'])]
    n = sum(1 for k in fx.files if k.endswith('_type'))
    return synth.clip(SYNTH_SEED[case], n, h, w)


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else float(10 * np.log10(255.0 ** 2 / mse))


def measure(case, precision, dev, codec=None):
    """-> dict of parity statistics of `precision` ('fp32' | 'bf16x3' | 'bf16') on fixture `case`."""
    from aivc_b200 import models, gop as G
    from aivc_b200.codec import FrameCodec, planes_to_device
    from aivc_b200.plan import Config
    fx = np.load(os.path.join(GOLDEN, 'cfg_%s.npz' % case))
    h, w = int(fx['H']), int(fx['W'])
    gop = G.generate_gop_struct(str(fx['gop']))
    names = sorted(gop, key=lambda f: int(f.split('_')[1]))
    order = G.coding_order(gop)
    clip = source_frames(case, fx)
    if codec is None:
        codec = FrameCodec(models.build_standin(**MODELS[case]), h, w, dev, Config(precision=precision))
    frames = {f: planes_to_device(clip[i], dev) for i, f in enumerate(names)}
    bts, rec = codec.encode_gop(frames, gop)
    dec = codec.decode_gop(bts, gop)              # leaves every latent's decoded symbols in its slot
    torch.cuda.synchronize(dev)
    out = {'case': case, 'precision': precision, 'closed_loop_exact': True, 'frames': {},
           'y_symbols': 0, 'y_mismatches': 0, 'y_max_abs_diff': 0, 'z_symbols': 0, 'z_mismatches': 0,
           'bytes': 0, 'oracle_bytes': 0, 'frames_bytes_identical': 0, 'frames_planes_identical': 0,
           'max_level_diff_subsampled': 0, 'max_abs_psnr_delta_db': 0.0}
    for i, f in enumerate(order):
        for a, b in zip(rec[f], dec[f]):
            if not torch.equal(a, b):
                out['closed_loop_exact'] = False
        fr = {}
        for net_name, eng in (('mof', codec.mof), ('codec', codec.codec)):
            key = '%s_%s_q' % (f, net_name)
            if key not in fx.files:
                continue
            q, z = (a.astype(np.int32) for a in codec.decoded_symbols(f, net_name))
            dq = q - fx[key].astype(np.int32)
            dz = z - fx['%s_%s_z' % (f, net_name)].astype(np.int32)
            fr[net_name] = {'y_mismatches': int((dq != 0).sum()), 'z_mismatches': int((dz != 0).sum())}
            out['y_symbols'] += q.size
            out['y_mismatches'] += int((dq != 0).sum())
            out['y_max_abs_diff'] = max(out['y_max_abs_diff'], int(np.abs(dq).max()))
            out['z_symbols'] += z.size
            out['z_mismatches'] += int((dz != 0).sum())
        planes = [p.cpu().numpy() for p in rec[f]]
        hc, wc = (h + 1) // 2, (w + 1) // 2
        shaped = [planes[0].reshape(h, w), planes[1].reshape(hc, wc), planes[2].reshape(hc, wc)]
        lev = max(int(np.abs(p[::4, ::4].astype(np.int32) - fx['%s_sub_%s' % (f, k)].astype(np.int32)).max())
                  for k, p in zip('yuv', shaped))
        src = clip[names.index(f)]
        ps = psnr(np.concatenate(planes), np.concatenate([p.reshape(-1) for p in src]))
        fr.update(bytes=len(bts[f]), oracle_bytes=int(fx[f + '_nbytes']),
                  bytes_identical=hashlib.md5(bts[f]).hexdigest() == str(fx[f + '_bytes_md5']),
                  planes_identical=hashlib.md5(b''.join(p.tobytes() for p in planes)).hexdigest() == str(fx[f + '_planes_md5']),
                  max_level_diff_subsampled=lev, psnr_vs_source_db=ps,
                  psnr_delta_db=ps - float(fx[f + '_psnr_vs_source']))
        out['frames'][f] = fr
        out['bytes'] += fr['bytes']
        out['oracle_bytes'] += fr['oracle_bytes']
        out['frames_bytes_identical'] += int(fr['bytes_identical'])
        out['frames_planes_identical'] += int(fr['planes_identical'])
        out['max_level_diff_subsampled'] = max(out['max_level_diff_subsampled'], lev)
        out['max_abs_psnr_delta_db'] = max(out['max_abs_psnr_delta_db'], abs(fr['psnr_delta_db']))
    out['n_frames'] = len(order)
    out['y_index_mismatch_rate'] = out['y_mismatches'] / max(out['y_symbols'], 1)
    out['bytes_delta'] = out['bytes'] - out['oracle_bytes']
    return out
