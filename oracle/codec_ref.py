"""ORACLE (test infrastructure only -- never imported by the product path).

CPU fp32 restatement of AIVC's per-frame encode / decode, composed from
``oracle.nn_ref`` and the C range coder ``oracle/torchac_ref.c``:

  decoder : real_life/decode.py:455-580 (Decoder.decode), :602-654 (CodecNetDecoder),
            :677-749 (MOFNetDecoder), :798-898 (ConditionalDecoder)
  framing : real_life/bitstream.py:186-304 (encode), :352-501 (decode)
  encoder : the missing models.FullNet.GOP_forward, reconstructed as the mirror of the
            decoder; its closed loop is what the decoder fixes.

Two integer-CDF modes for the y latents:
  'reference' -- torch fp32 Laplace table [C,h,w,514] exactly as bitstream.py:127-154,
                 then torchac's normalisation (small sizes only: the table is 2 KB/symbol);
  'spec'      -- the deterministic evaluation of oracle/laplace_spec.c that the product
                 implements on the device and in its host decoder.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

from . import nn_ref as R

FRAME_I, FRAME_P, FRAME_B = 0, 1, 2
AC_MAX_VAL = 256
LP = 2 * AC_MAX_VAL + 2

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, '_build', 'liboracle.so')
        if not os.path.exists(so):
            subprocess.check_call(['make', '-C', _HERE], stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(so)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.tac_ref_encode_table.restype = ctypes.c_size_t
        L.tac_ref_encode_table.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_size_t, ctypes.POINTER(u8p)]
        L.tac_ref_encode_bounds.restype = ctypes.c_size_t
        L.tac_ref_encode_bounds.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                            ctypes.POINTER(u8p)]
        L.tac_ref_free.argtypes = [u8p]
        L.tac_ref_decode_table.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        L.laplace_spec_table.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        L.laplace_spec_bounds.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                          ctypes.c_void_p, ctypes.c_void_p]
        L.laplace_spec_cdf_int.restype = ctypes.c_uint32
        L.laplace_spec_cdf_int.argtypes = [ctypes.c_float, ctypes.c_int]
        _LIB = L
    return _LIB


# ------------------------------------------------------------------ range coder wrappers
def rc_encode_table(cdf_u16, sym_i16):
    """cdf_u16: [n, Lp] uint16, sym_i16: [n] int16 -> bytes."""
    cdf = np.ascontiguousarray(cdf_u16, dtype=np.uint16)
    sym = np.ascontiguousarray(sym_i16, dtype=np.int16)
    out = ctypes.POINTER(ctypes.c_uint8)()
    n = lib().tac_ref_encode_table(cdf.ctypes.data, cdf.shape[-1], sym.ctypes.data, sym.size,
                                   ctypes.byref(out))
    data = ctypes.string_at(out, n)
    lib().tac_ref_free(out)
    return data


def rc_encode_bounds(c_low, c_high):
    lo = np.ascontiguousarray(c_low, dtype=np.uint32)
    hi = np.ascontiguousarray(c_high, dtype=np.uint32)
    out = ctypes.POINTER(ctypes.c_uint8)()
    n = lib().tac_ref_encode_bounds(lo.ctypes.data, hi.ctypes.data, lo.size, ctypes.byref(out))
    data = ctypes.string_at(out, n)
    lib().tac_ref_free(out)
    return data


def rc_decode_table(cdf_u16, data, n):
    cdf = np.ascontiguousarray(cdf_u16, dtype=np.uint16)
    sym = np.empty(n, dtype=np.int16)
    buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
    lib().tac_ref_decode_table(cdf.ctypes.data, cdf.shape[-1], buf.ctypes.data, len(data),
                               sym.ctypes.data, n)
    return sym


def laplace_table_spec(sigma_flat):
    s = np.ascontiguousarray(sigma_flat, dtype=np.float32)
    out = np.empty((s.size, LP), dtype=np.uint16)
    lib().laplace_spec_table(s.ctypes.data, s.size, out.ctypes.data)
    return out


def laplace_table_reference(sigma_flat):
    t = R.laplace_cdf_table(torch.as_tensor(np.asarray(sigma_flat, dtype=np.float32)))
    return R.cdf_float_to_int(t).numpy().astype(np.uint16)


def z_table_u16(pdf_z):
    """[C_z, 514] uint16 (bitstream.py:82-125 + torchac normalisation)."""
    with torch.no_grad():
        return R.cdf_float_to_int(R.z_cdf_table(pdf_z, AC_MAX_VAL)).numpy().astype(np.uint16)


# ------------------------------------------------------------------ per-latent framing
def ac_encode_latent(x, mode, sigma=None, z_table=None, cdf_mode='spec', first_of_i_frame=False):
    """bitstream.py:186-304 without the file system: returns the bytes appended to the
    frame's bitstream for one latent.  x: integral-valued float tensor [1,C,H,W]."""
    body = b''
    if mode == 'laplace':
        nz = (x.abs().sum(dim=(2, 3)).squeeze(0) != 0)
        idx = [i for i in range(nz.numel()) if bool(nz[i])]
        body += len(idx).to_bytes(1, 'big') + bytes(idx)
        if idx:
            xs = x[0, idx].reshape(-1)
            sg = sigma[0, idx].reshape(-1).numpy()
            table = laplace_table_spec(sg) if cdf_mode == 'spec' else laplace_table_reference(sg)
            sym = (xs + AC_MAX_VAL).to(torch.int16).numpy()
            body += rc_encode_table(table, sym)
    else:
        _, c, h, w = x.shape
        table = np.repeat(z_table, h * w, axis=0)                 # NCHW order: channel-major
        sym = (x.reshape(-1) + AC_MAX_VAL).to(torch.int16).numpy()
        body += rc_encode_table(table, sym)
    out = len(body).to_bytes(4, 'big') + body
    if first_of_i_frame:
        # bitstream.py:292-296: an I frame has no MOFNet part -> two empty sections first
        out = (0).to_bytes(4, 'big') + (0).to_bytes(4, 'big') + out
    return out


def split_frame_sections(frame_bytes):
    """bitstream.py:394-416: four [4-byte length][payload] sections per frame."""
    secs, pos = [], 0
    for _ in range(4):
        n = int.from_bytes(frame_bytes[pos:pos + 4], 'big')
        secs.append(frame_bytes[pos + 4:pos + 4 + n])
        pos += 4 + n
    return secs


def ac_decode_latent(sec, mode, shape, sigma=None, z_table=None, cdf_mode='spec'):
    """bitstream.py:352-501 on one already-isolated section."""
    _, c, h, w = shape
    if mode == 'laplace':
        n_sent = sec[0]
        out = torch.zeros(shape)
        if n_sent:
            idx = list(sec[1:1 + n_sent])
            sg = sigma[0, idx].reshape(-1).numpy()
            table = laplace_table_spec(sg) if cdf_mode == 'spec' else laplace_table_reference(sg)
            sym = rc_decode_table(table, sec[1 + n_sent:], len(idx) * h * w)
            out[0, idx] = torch.from_numpy(sym.astype(np.float32) - AC_MAX_VAL).view(len(idx), h, w)
        return out
    table = np.repeat(z_table, h * w, axis=0)
    sym = rc_decode_table(table, sec, c * h * w)
    return torch.from_numpy(sym.astype(np.float32) - AC_MAX_VAL).view(shape)


# ------------------------------------------------------------------ conditional coder
def _gain(net, frame_type):
    if not net.flag_gain_p_b or frame_type == FRAME_I:
        return net.gain_I
    return net.gain_P if frame_type == FRAME_P else net.gain_B


def latent_dims(h, w):
    """y and z sizes: four ceil-halvings for y, two more for z (header.py:74-80 stores them)."""
    c = lambda v, n: v if n == 0 else c((v + 1) // 2, n - 1)
    return (c(h, 4), c(w, 4)), (c(h, 6), c(w, 6))


def cond_encode(net, x_in, in_shortcut, frame_type, z_table, idx_rate=0., cdf_mode='spec',
                first_of_i_frame=False):
    """Analysis side of a ConditionalNet + its own decoding (closed loop).
    Returns (bytes, x_hat_raw, aux)."""
    g = _gain(net, frame_type)
    y = R.forward_module(net.g_a, x_in) * R.gain_vector(g, idx_rate, 'enc')
    z_hat = torch.round(R.forward_module(net.h_a, y))
    z_hat = torch.clamp(z_hat, -AC_MAX_VAL, AC_MAX_VAL - 1)
    h_y, w_y = y.shape[2:]
    mu, sigma = R.mu_sigma(R.forward_module(net.h_s, z_hat)[:, :, :h_y, :w_y], net.nb_ft_y)
    q = torch.clamp(torch.round(y - mu), -AC_MAX_VAL, AC_MAX_VAL - 1)
    data = ac_encode_latent(z_hat, 'pmf', z_table=z_table, first_of_i_frame=first_of_i_frame)
    data += ac_encode_latent(q, 'laplace', sigma=sigma, cdf_mode=cdf_mode)
    x_hat = cond_synthesis(net, q, mu, in_shortcut, frame_type, idx_rate)
    return data, x_hat, {'y': y, 'z_hat': z_hat, 'q': q, 'mu': mu, 'sigma': sigma}


def cond_synthesis(net, q, mu, in_shortcut, frame_type, idx_rate=0.):
    """decode.py:867-896."""
    y_hat = (q + mu) * R.gain_vector(_gain(net, frame_type), idx_rate, 'dec')
    if in_shortcut is not None and getattr(net, 'g_a_ref', None) is not None:
        sc = R.forward_module(net.g_a_ref, in_shortcut)
    else:
        sc = torch.zeros((1, net.out_c_shortcut_y, q.shape[2], q.shape[3]))
    return R.forward_module(net.g_s, torch.cat((y_hat, sc), dim=1))


def cond_decode(net, sec_z, sec_y, in_shortcut, frame_type, dims_y, dims_z, z_table,
                idx_rate=0., cdf_mode='spec'):
    """ConditionalDecoder.decode -- decode.py:798-898."""
    z_hat = ac_decode_latent(sec_z, 'pmf', (1, net.nb_ft_z) + tuple(dims_z), z_table=z_table)
    h_y, w_y = dims_y
    mu, sigma = R.mu_sigma(R.forward_module(net.h_s, z_hat)[:, :, :h_y, :w_y], net.nb_ft_y)
    q = ac_decode_latent(sec_y, 'laplace', (1, net.nb_ft_y, h_y, w_y), sigma=sigma,
                         cdf_mode=cdf_mode)
    return cond_synthesis(net, q, mu, in_shortcut, frame_type, idx_rate), \
        {'z_hat': z_hat, 'q': q, 'mu': mu, 'sigma': sigma}


# ------------------------------------------------------------------ frames
def zero_yuv(h, w):
    return {'y': torch.zeros(1, 1, h, w), 'u': torch.zeros(1, 1, (h + 1) // 2, (w + 1) // 2),
            'v': torch.zeros(1, 1, (h + 1) // 2, (w + 1) // 2)}


class Tables:
    def __init__(self, model):
        self.mof = z_table_u16(model.mode_net.mode_net.pdf_z)
        self.codec = z_table_u16(model.codec_net.codec_net.pdf_z)


def encode_frame(model, tables, frame, prev_dic, next_dic, frame_type, idx_rate=0.,
                 cdf_mode='spec'):
    """One frame through MOFNet + motion compensation + CodecNet, closed loop.
    Returns (frame_bytes, reconstructed 8-bit-levelled YUV420 dict, aux)."""
    with torch.no_grad():
        h, w = frame['y'].shape[2:]
        code = R.input_layer(frame)
        prev_ref, next_ref = R.input_layer(prev_dic), R.input_layer(next_dic)
        data, aux = b'', {}
        if frame_type == FRAME_I:
            alpha = torch.ones(1, 3, h, w)
            x_warp = torch.zeros(1, 3, h, w)
        else:
            mnet = model.mode_net.mode_net
            sc = torch.cat((prev_ref, next_ref), 1) if frame_type == FRAME_B else None
            d, raw, a = cond_encode(mnet, torch.cat((code, prev_ref, next_ref), 1), sc, frame_type,
                                    tables.mof, idx_rate, cdf_mode)
            data += d
            aux['mof'] = a
            alpha, beta, v_prev, v_next = R.mofnet_post(raw, h, w, frame_type == FRAME_P)
            x_warp = R.motion_compensation(prev_ref, next_ref, v_prev, v_next, beta)
        skip = (1 - alpha) * x_warp
        pred = alpha * x_warp
        cnet = model.codec_net.codec_net
        d, raw, a = cond_encode(cnet, torch.cat((code, pred), 1),
                                pred if frame_type != FRAME_I else None, frame_type, tables.codec,
                                idx_rate, cdf_mode, first_of_i_frame=(frame_type == FRAME_I))
        data += d
        aux['codec'] = a
        rec = R.finalize_frame(raw[:, :, :h, :w] + skip, h, w)
        aux.update(alpha=alpha, x_warp=x_warp)
        return data, rec, aux


def decode_frame(model, tables, frame_bytes, prev_dic, next_dic, frame_type, h, w, idx_rate=0.,
                 cdf_mode='spec'):
    """Decoder.decode -- decode.py:455-580."""
    with torch.no_grad():
        dims_y, dims_z = latent_dims(h, w)
        secs = split_frame_sections(frame_bytes)
        prev_ref, next_ref = R.input_layer(prev_dic), R.input_layer(next_dic)
        aux = {}
        if frame_type == FRAME_I:
            alpha = torch.ones(1, 3, h, w)
            x_warp = torch.zeros(1, 3, h, w)
        else:
            sc = torch.cat((prev_ref, next_ref), 1) if frame_type == FRAME_B else None
            raw, aux['mof'] = cond_decode(model.mode_net.mode_net, secs[0], secs[1], sc,
                                          frame_type, dims_y, dims_z, tables.mof, idx_rate, cdf_mode)
            alpha, beta, v_prev, v_next = R.mofnet_post(raw, h, w, frame_type == FRAME_P)
            x_warp = R.motion_compensation(prev_ref, next_ref, v_prev, v_next, beta)
        skip = (1 - alpha) * x_warp
        pred = alpha * x_warp
        raw, aux['codec'] = cond_decode(model.codec_net.codec_net, secs[2], secs[3],
                                        pred if frame_type != FRAME_I else None, frame_type,
                                        dims_y, dims_z, tables.codec, idx_rate, cdf_mode)
        return R.finalize_frame(raw[:, :, :h, :w] + skip, h, w), aux


def _coding_order(gop):
    return sorted(gop, key=lambda f: gop[f]['coding_order'])


def encode_gop(model, tables, frames, gop, idx_rate=0., cdf_mode='spec'):
    """frames: {'frame_i': yuv dict}; gop: GOP_structure dict. Returns
    ({'frame_i': bytes}, {'frame_i': recon})."""
    h, w = frames['frame_0']['y'].shape[2:]
    out_b, rec = {}, {}
    for f in _coding_order(gop):
        t = gop[f]['type']
        prev = rec[gop[f]['prev_ref']] if t != FRAME_I else zero_yuv(h, w)
        nxt = rec[gop[f]['next_ref']] if t == FRAME_B else zero_yuv(h, w)
        out_b[f], rec[f], _ = encode_frame(model, tables, frames[f], prev, nxt, t, idx_rate, cdf_mode)
    return out_b, rec


def decode_gop(model, tables, frame_bytes, gop, h, w, idx_rate=0., cdf_mode='spec'):
    """decode_one_GOP -- decode.py:193-326 (frames in coding order, zero refs when absent)."""
    rec = {}
    for f in _coding_order(gop):
        t = gop[f]['type']
        prev = rec[gop[f]['prev_ref']] if t != FRAME_I else zero_yuv(h, w)
        nxt = rec[gop[f]['next_ref']] if t == FRAME_B else zero_yuv(h, w)
        rec[f], _ = decode_frame(model, tables, frame_bytes[f], prev, nxt, t, h, w, idx_rate,
                                 cdf_mode)
    return rec
