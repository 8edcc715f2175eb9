"""Test helper: the reference's DECODER composition, restated call for call from real_life/decode.py:455-898
(Decoder.decode -> MOFNetDecoder.decode / CodecNetDecoder.decode -> ConditionalDecoder.decode), but executed through the
drop-in nn.Module classes' own forward() -- net.g_s(x), net.h_s(x), net.g_a_ref(x), net.pdf_parameterizer(x),
net.gain_X({...}), model.in_layer(dic), model.motion_compensation({...}), model.out_layer(x) -- on CUDA tensors, the way
the reference's Decoder would drive them.  Entropy decoding goes through aivc_b200.entropy (the torchac replacement).
This is the path a user gets who lets the reference's own decoder run on the mirrors, as opposed to the fused FrameCodec."""
import numpy as np
import torch

from aivc_b200 import entropy
from aivc_b200.codec import latent_dims

FRAME_I, FRAME_P, FRAME_B = 0, 1, 2


def _gain(net, ft):
    if not net.flag_gain_p_b or ft == FRAME_I:
        return net.gain_I
    return net.gain_P if ft == FRAME_P else net.gain_B


def cond_decode(net, sec_z, sec_y, in_shortcut, ft, dims_y, dims_z, dev, idx_rate=0.):
    """ConditionalDecoder.decode -- decode.py:798-898"""
    (hy, wy), (hz, wz) = dims_y, dims_z
    table = entropy.z_table_u16(net.pdf_z)
    z = entropy.decode_z(table, sec_z, net.nb_ft_z, hz, wz)                                   # :844-850
    z_hat = torch.from_numpy(z.astype(np.float32))[None].to(dev)
    prm = net.pdf_parameterizer(net.h_s(z_hat)[:, :, :hy, :wy].contiguous())[0]               # :853-855
    b = (prm['sigma'][0] / torch.sqrt(torch.tensor([2.0], device=dev))).cpu().numpy()         # bitstream.py:141
    q = entropy.decode_y(sec_y, np.ascontiguousarray(b), net.nb_ft_y, hy, wy)                 # :858-867
    y_hat = torch.from_numpy(q.astype(np.float32))[None].to(dev) + prm['mu']
    y_hat = _gain(net, ft)({'x': y_hat, 'idx_rate': idx_rate, 'mode': 'dec'})['output']       # :870-885
    if in_shortcut is not None and getattr(net, 'g_a_ref', None) is not None:                 # :888-892
        sc = net.g_a_ref(in_shortcut)
    else:
        sc = torch.zeros((1, net.out_c_shortcut_y, hy, wy), device=dev)
    return net.g_s(torch.cat((y_hat, sc), dim=1))                                             # :895-896


def decode_frame(model, frame_bytes, prev_dic, next_dic, ft, h, w, dev):
    """Decoder.decode -- decode.py:455-580.  prev_dic / next_dic: YUV420 dicts of CUDA tensors in [0, 1]."""
    dims_y, dims_z = latent_dims(h, w)
    secs = entropy.split_sections(frame_bytes)
    prev_ref, next_ref = model.in_layer(prev_dic), model.in_layer(next_dic)                   # :493-494
    if ft == FRAME_I:                                                                         # :500-504
        alpha = torch.ones((1, 3, h, w), device=dev)
        x_warp = torch.zeros((1, 3, h, w), device=dev)
    else:
        sc = torch.cat((prev_ref, next_ref), 1) if ft == FRAME_B else None                    # :710-714
        raw = cond_decode(model.mode_net.mode_net, secs[0], secs[1], sc, ft, dims_y, dims_z, dev)[:, :, :h, :w]
        alpha = torch.clamp(raw[:, 0:1] + 0.5, 0., 1.).repeat(1, 3, 1, 1)                     # :731-739
        beta = torch.clamp(raw[:, 1:2] + 0.5, 0., 1.).repeat(1, 3, 1, 1)
        v_prev, v_next = raw[:, 2:4].contiguous(), raw[:, 4:6].contiguous()
        if ft == FRAME_P:
            beta, v_next = torch.ones_like(beta), torch.zeros_like(v_next)
        x_warp = model.motion_compensation({'prev': prev_ref, 'next': next_ref, 'v_prev': v_prev, 'v_next': v_next,
                                            'beta': beta, 'interpol_mode': 'bilinear'})['x_warp']     # :524-533
    skip = (1 - alpha) * x_warp                                                               # :536
    pred = alpha * x_warp
    raw = cond_decode(model.codec_net.codec_net, secs[2], secs[3], pred if ft != FRAME_I else None, ft, dims_y, dims_z, dev)
    out = model.out_layer(raw[:, :, :h, :w].contiguous() + skip)                              # :549-553
    hc, wc = (h + 1) // 2, (w + 1) // 2
    for k in 'uv':                                                                            # :557-571 (replicate-pad odd sizes)
        t = out[k]
        if t.shape[2] < hc or t.shape[3] < wc:
            t = torch.nn.functional.pad(t, (0, wc - t.shape[3], 0, hc - t.shape[2]), mode='replicate')
        out[k] = t
    return {k: torch.round(255. * torch.clamp(v, 0., 1.)) / 255. for k, v in out.items()}     # img_processing.py:68-73


def decode_gop(model, frame_bytes, gop, h, w, dev):
    hc, wc = (h + 1) // 2, (w + 1) // 2
    zero = lambda: {'y': torch.zeros(1, 1, h, w, device=dev), 'u': torch.zeros(1, 1, hc, wc, device=dev),
                    'v': torch.zeros(1, 1, hc, wc, device=dev)}
    rec = {}
    for f in sorted(gop, key=lambda f: gop[f]['coding_order']):
        t = gop[f]['type']
        prev = rec[gop[f]['prev_ref']] if t != FRAME_I else zero()
        nxt = rec[gop[f]['next_ref']] if t == FRAME_B else zero()
        rec[f] = decode_frame(model, frame_bytes[f], prev, nxt, t, h, w, dev)
    return rec
