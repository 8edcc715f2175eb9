"""Reference-facing entry points on top of FrameCodec.

``gop_forward(model, model_input)`` implements the contract of the (missing) upstream
``models.FullNet.GOP_forward`` as read off its callers:
  input  : model_management.py:307-317  (GOP_struct, GOP_struct_name, raw_frames, idx_rate,
           index_GOP_in_video, generate_bitstream, real_idx_first_frame, bitstream_dir,
           flag_bitstream_debug)
  output : net_out['frame_i'] with x_hat (YUV420 dict), alpha, beta, warping, code and the four
           rate tensors consumed by loss_function.py:158-204
  files  : with generate_bitstream it leaves '<bitstream_dir>/<idx_gop>g' (what cat_one_gop,
           cat_binary_files.py:19-101, would have produced) and 'data_dim.pkl' (header.py:175-177),
           so the reference's cat_one_video (model_management.py:216-223) finishes the job.

``decode_video(model, video_bytes)`` is decode_one_video (real_life/decode.py:44-154) without
the temp-dir round trips: bytes in, decoded uint8 planes out.
"""
import collections
import os
import pickle
import weakref

import numpy as np
import torch

from . import container
from .codec import FrameCodec, latent_dims
from .gop import generate_gop_struct, FRAME_I
from .plan import Config

_CODECS = weakref.WeakKeyDictionary()        # model -> OrderedDict(key -> (weights version, FrameCodec)), LRU
MAX_CODECS_PER_MODEL = 2                     # a 1080p codec holds 3-6 GB of plans and buffers


def codec_for(model, h, w, device, cfg=None, idx_rate=0.):
    """FrameCodec of `model` for this frame size / device / config / rate index.  Cached per model object in a
    WeakKeyDictionary (not in `model.__dict__`: whole-module pickles and deepcopies stay clean; not by id(model): a
    dead model's codec is never handed to a new model at the same address), rebuilt when the model's weights have
    changed since they were packed, at most MAX_CODECS_PER_MODEL alive per model (least recently used goes)."""
    from .plan import weights_version
    cfg = cfg or Config()
    cache = _CODECS.setdefault(model, collections.OrderedDict())
    key = (h, w, str(device), cfg.key(), float(idx_rate))
    ver = weights_version(model)
    hit = cache.get(key)
    if hit is None or hit[0] != ver:
        cache.pop(key, None)
        while len(cache) >= MAX_CODECS_PER_MODEL:
            cache.popitem(last=False)
        hit = cache[key] = (ver, FrameCodec(model, h, w, device, cfg, idx_rate))
    cache.move_to_end(key)
    return hit[1]


def _to_planes(yuv, device):
    """{'y','u','v'} fp32 [1,1,H,W] in [0,1] (any device) -> flat uint8 device planes."""
    return tuple((yuv[k].to(device).float().clamp(0, 1) * 255.).round().to(torch.uint8).reshape(-1).contiguous()
                 for k in 'yuv')


def _to_yuv(planes, h, w):
    hc, wc = (h + 1) // 2, (w + 1) // 2
    y, u, v = planes
    return {'y': y.view(1, 1, h, w).float() / 255., 'u': u.view(1, 1, hc, wc).float() / 255.,
            'v': v.view(1, 1, hc, wc).float() / 255.}


def _section_bits(frame_bytes):
    """Real coded size of the four sections of a frame, in bits (length prefixes excluded)."""
    bits, pos = [], 0
    for _ in range(4):
        n = int.from_bytes(frame_bytes[pos:pos + 4], 'big')
        bits.append(8.0 * n)
        pos += 4 + n
    return bits


def _rate_z(cond_net, z_i16, device):
    """Per-symbol rate estimate of the z latent, [1, C_z, h_z, w_z] bits: EntropyCoder(BallePdfEstim(z_hat))
    (pdf_estimator.py:172-202, entropy_coder.py:25-30).  A few thousand symbols through a 1-3-3-3-1 MLP per
    channel: evaluated by the (host-side) Balle mirror, like the z table."""
    z = torch.from_numpy(z_i16.astype(np.float32))[None]
    with torch.no_grad():
        p = cond_net.pdf_z(z.to(next(cond_net.pdf_z.parameters()).device))
    return (-torch.log2(torch.clamp(p.float(), 2.0 ** -16, 1.0))).to(device)


def gop_forward(model, model_input, device=None, cfg=None):
    """FullNet.GOP_forward (see the module docstring).  net_out[frame] holds, as the reference's encoder does:
      x_hat                      reconstructed YUV420 dict (8-bit levels / 255)
      code                       in_layer(raw frame), [1, 3, H, W]
      alpha, beta                clamp(MOFNet channel 0 / 1 + 0.5, 0, 1) repeated to 3 channels (decode.py:731-739);
                                 I frames: ones (decode.py:500-504); P frames: beta = 1
      warping                    x_warp = beta w(prev, v_prev) + (1 - beta) w(next, v_next); I frames: zeros
      mode_rate_y/z, codec_rate_y/z   per-symbol rate ESTIMATES in bits, [1, C, h, w] (pdf_estimator.py:27-70,
                                 172-202; entropy_coder.py:25-30) -- what loss_function.py:158,186 sums; I frames
                                 have no MOFNet latents: zeros
      coded_bits                 (extra key) the REAL coded size of the four sections, from the bitstream"""
    gop = model_input['GOP_struct']
    name = model_input.get('GOP_struct_name') or ''
    raw = model_input['raw_frames']
    idx_rate = float(model_input.get('idx_rate', 0.) or 0.)
    first = raw['frame_0']['y']
    h, w = first.shape[2:]
    if device is None:
        device = first.device if first.is_cuda else torch.device('cuda', torch.cuda.current_device())
    device = torch.device(device)
    codec = codec_for(model, h, w, device, cfg, idx_rate)
    frames = {f: _to_planes(raw[f], device) for f in gop}
    aux = {}
    bts, rec = codec.encode_gop(frames, gop, aux=aux)

    net_out = {}
    in_layer = model.in_layer
    mnet, cnet = model.mode_net.mode_net, model.codec_net.codec_net
    (hy, wy), _ = latent_dims(h, w)
    for f in gop:
        a = aux[f]
        x_hat = _to_yuv(rec[f], h, w)
        code = in_layer({k: (frames[f][i].view(1, 1, *x_hat[k].shape[2:]).float() / 255.)
                         for i, k in enumerate('yuv')})
        out = {'x_hat': x_hat, 'code': code,
               'codec_rate_y': a['codec_rate_y'].view(1, cnet.nb_ft_y, hy, wy),
               'codec_rate_z': _rate_z(cnet, a['codec_keep']['z'], device),
               'coded_bits': dict(zip(('mode_z', 'mode_y', 'codec_z', 'codec_y'), _section_bits(bts[f])))}
        if gop[f]['type'] == FRAME_I:
            one = torch.ones((1, 3, h, w), device=device)
            out.update(alpha=one, beta=one, warping=torch.zeros((1, 3, h, w), device=device),
                       mode_rate_y=torch.zeros((1, mnet.nb_ft_y, hy, wy), device=device),
                       mode_rate_z=torch.zeros_like(out['codec_rate_z']))
        else:
            wa = a['warp'].view(5, h, w)
            out.update(alpha=wa[0].expand(1, 3, h, w), beta=wa[1].expand(1, 3, h, w), warping=wa[2:5].unsqueeze(0),
                       mode_rate_y=a['mode_rate_y'].view(1, mnet.nb_ft_y, hy, wy),
                       mode_rate_z=_rate_z(mnet, a['mode_keep']['z'], device))
        net_out[f] = out

    if model_input.get('generate_bitstream'):
        d = model_input.get('bitstream_dir') or './'
        if not d.endswith('/'):
            d += '/'
        os.makedirs(d, exist_ok=True)
        order = sorted(gop, key=lambda f: int(f.split('_')[1]))
        idx_gop = int(model_input.get('index_GOP_in_video', 0) or 0)
        with open(d + str(idx_gop) + 'g', 'wb') as fo:
            fo.write(container.pack_gop(name, [bts[f] for f in order], idx_rate))
        dims_y, dims_z = latent_dims(h, w)
        with open(d + 'data_dim.pkl', 'wb') as fo:
            pickle.dump({'x': (h, w), 'y': dims_y, 'z': dims_z}, fo, pickle.HIGHEST_PROTOCOL)
    return net_out


def compute_metrics_one_gop(net_out, target, nb_pad_frame=0):
    """The logging half of compute_metrics_one_GOP (loss_function.py:103-257) on the device metrics kernel
    (csrc/metrics.cu): per frame mse, psnr, ms_ssim, ms_ssim_db, mse_warping, psnr_warping, the three rates in bpp,
    mean_alpha, mean_beta, h, w; plus the 'GOP' average (distortions over the real frames, rates over all --
    average_N_frame, loss_function.py:260-339).  target: {'frame_i': YUV420 dict in [0, 1]}."""
    from . import metrics
    result = {}
    for f in target:
        o = net_out[f]
        h, w = target[f]['y'].shape[2:]
        dev = o['code'].device
        m = metrics.frame_metrics(_to_planes(o['x_hat'], dev), _to_planes(target[f], dev), h, w)
        npx = float(h * w)
        r = {k: float(o[k].sum()) / npx for k in ('mode_rate_y', 'mode_rate_z', 'codec_rate_y', 'codec_rate_z')}
        mse_w = float(((o['warping'] - o['code']) ** 2).mean())
        result[f] = {'mse': m['mse'], 'psnr': m['psnr'], 'ms_ssim': m['ms_ssim'], 'ms_ssim_db': m['ms_ssim_db'],
                     'mse_warping': mse_w, 'psnr_warping': 10 * np.log10(1. / max(mse_w, 1e-20)),
                     'codec_rate_bpp': r['codec_rate_y'] + r['codec_rate_z'],
                     'mode_rate_bpp': r['mode_rate_y'] + r['mode_rate_z'],
                     'total_rate_bpp': sum(r.values()),
                     'mean_alpha': float(o['alpha'].mean()), 'mean_beta': float(o['beta'].mean()), 'h': float(h), 'w': float(w)}
    names = sorted(result, key=lambda f: int(f.split('_')[1]))
    real = names[:len(names) - nb_pad_frame] if nb_pad_frame else names
    gop = {}
    for k in result[names[0]]:       # distortions over the real frames only, everything else over all frames
        src = real if k in ('mse', 'mse_warping', 'ms_ssim', 'ms_ssim_db', 'psnr') else names
        gop[k] = float(np.mean([result[f][k] for f in src]))
    gop['psnr'] = 10 * np.log10(1. / max(gop['mse'], 1e-20))                 # of the average mse, not the average psnr
    gop['psnr_warping'] = 10 * np.log10(1. / max(gop['mse_warping'], 1e-20))
    gop['ms_ssim_db'] = -10.0 * np.log10(max(1. - gop['ms_ssim'], 1e-20))
    result['GOP'] = gop
    return result


def encode_video(model, gops_of_frames, gop_name, h, w, device='cuda:0', cfg=None, idx_rate=0., idx_first=0):
    """gops_of_frames: list (one per GOP) of {'frame_i': (y,u,v) uint8 device planes}.
    Returns the complete .bin byte string (same layout as cat_one_video writes)."""
    gop = generate_gop_struct(gop_name)
    order = sorted(gop, key=lambda f: int(f.split('_')[1]))
    codec = codec_for(model, h, w, torch.device(device), cfg, idx_rate)
    packed = []
    for frames in gops_of_frames:
        bts, _ = codec.encode_gop(frames, gop)
        packed.append(container.pack_gop(gop_name, [bts[f] for f in order], idx_rate))
    dims_y, dims_z = latent_dims(h, w)
    n = len(order) * len(gops_of_frames)
    return container.pack_video((h, w), dims_y, dims_z, packed, idx_first, idx_first + n - 1)


def decode_video(model, video_bytes, device='cuda:0', cfg=None):
    """-> (list over GOPs of {'frame_i': (y,u,v) uint8 device planes}, data_dim, idx_first, idx_last)"""
    dims, gops, first, last = container.unpack_video(video_bytes)
    h, w = dims['x']
    out = []
    for g in gops:
        name, idx_rate, frames = container.unpack_gop(g)
        gop = generate_gop_struct(name)
        order = sorted(gop, key=lambda f: int(f.split('_')[1]))
        if len(order) != len(frames):
            raise ValueError('GOP %s announces %d frames, bitstream holds %d' % (name, len(order), len(frames)))
        codec = codec_for(model, h, w, torch.device(device), cfg, idx_rate)
        out.append(codec.decode_gop(dict(zip(order, frames)), gop))
    return out, dims, first, last


def encode_yuv(model, path_in, gop_name, idx_start=0, idx_end=-1, device='cuda:0', cfg=None, idx_rate=0.):
    """Planar .yuv file -> AIVC video bitstream bytes, without the reference's PNG detour (yuvio.py)."""
    from . import yuvio
    return yuvio.encode_yuv_file(codec_for, model, path_in, gop_name, idx_start, idx_end, device, cfg, idx_rate)


def decode_to_yuv(model, video_bytes, path_out, device='cuda:0', cfg=None):
    """AIVC video bitstream bytes -> planar .yuv file; returns the number of frames written."""
    from . import yuvio
    return yuvio.decode_to_yuv_file(decode_video, model, video_bytes, path_out, device, cfg)
