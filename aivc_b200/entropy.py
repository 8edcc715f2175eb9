"""Host side of the entropy-coding hand-off: framing of one latent and calls into the
C++ range coder (aivc_b200/csrc/rangecoder.cpp).

Mirrors ``ArithmeticCoder.encode/decode`` (real_life/bitstream.py:186-304, 352-501) without the
file system, the [C,H,W,514] float CDF tables, or the debug re-decode: the device hands over
16-bit CDF bounds per symbol (encoder) or the Laplace scale per symbol (decoder).

Per-latent section:  [4-byte big-endian length][laplace only: 1 byte n_ch, n_ch channel ids][payload]
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

AC_MAX_VAL = 256
LP = 2 * AC_MAX_VAL + 2


def z_table_u16(pdf_z):
    """[C_z, 514] uint16 CDF table of the factorised prior, computed once per model on the
    host (bitstream.py:82-125 + torchac's normalisation round(cdf*(2^16-513)) + i mod 2^16)."""
    with torch.no_grad():
        idx = (torch.arange(LP, dtype=torch.float32) - AC_MAX_VAL - 0.5)
        idx = idx.view(1, 1, -1, 1).repeat(1, pdf_z.nb_channel, 1, 1)
        p = next(pdf_z.parameters())
        cdf = pdf_z.cdf(idx.to(p.device)).squeeze(-1).squeeze(0).float().cpu()
        scaled = (cdf * float(65536 - (LP - 1))).round().to(torch.int64)
        tab = (scaled + torch.arange(LP, dtype=torch.int64)) & 0xFFFF
    return np.ascontiguousarray(tab.numpy().astype(np.uint16))


def _encode_bounds(bounds_u32):
    L = _lib.lib()
    out = np.empty(L.aivc_rc_bound(bounds_u32.size), dtype=np.uint8)
    n = C.c_size_t()
    _lib.check(L.aivc_rc_encode_bounds(bounds_u32.ctypes.data, bounds_u32.size, out.ctypes.data,
                                       out.size, C.byref(n)))
    return out[:n.value].tobytes()


def encode_z(table, z_i16):
    """z_i16: [C, h, w] int16 (host) -> section bytes."""
    L = _lib.lib()
    c = z_i16.shape[0]
    hw = z_i16.size // c
    out = np.empty(L.aivc_rc_bound(z_i16.size), dtype=np.uint8)
    n = C.c_size_t()
    z = np.ascontiguousarray(z_i16)
    _lib.check(L.aivc_rc_encode_table(table.ctypes.data, z.ctypes.data, c, hw, out.ctypes.data,
                                      out.size, C.byref(n)))
    body = out[:n.value].tobytes()
    return len(body).to_bytes(4, 'big') + body


def encode_y(bounds_u32, nz_i32):
    """bounds: [C, h, w] uint32 (host), nz: [C] int32 flags -> section bytes.
    Only channels with a non-zero symbol are coded (bitstream.py:241-255)."""
    idx = np.nonzero(nz_i32)[0]
    body = len(idx).to_bytes(1, 'big') + bytes(int(i) for i in idx)
    if len(idx):
        sel = np.ascontiguousarray(bounds_u32[idx]).reshape(-1)
        body += _encode_bounds(sel)
    return len(body).to_bytes(4, 'big') + body


def split_sections(frame_bytes):
    """The four [length][payload] sections of a frame: mofnet_z, mofnet_y, codecnet_z,
    codecnet_y (bitstream.py:22-56, 394-416)."""
    secs, pos = [], 0
    for _ in range(4):
        if pos + 4 > len(frame_bytes):
            raise ValueError('truncated frame bitstream')
        n = int.from_bytes(frame_bytes[pos:pos + 4], 'big')
        secs.append(frame_bytes[pos + 4:pos + 4 + n])
        pos += 4 + n
    return secs


def decode_z(table, sec, c, h, w):
    L = _lib.lib()
    out = np.empty((c, h, w), dtype=np.int16)
    buf = np.frombuffer(sec, dtype=np.uint8) if len(sec) else np.zeros(1, np.uint8)
    _lib.check(L.aivc_rc_decode_table(table.ctypes.data, buf.ctypes.data, len(sec), c, h * w,
                                      out.ctypes.data))
    return out


def y_channels(sec):
    """Channel indices signalled in a 'laplace' section, and the payload."""
    n = sec[0]
    return list(sec[1:1 + n]), sec[1 + n:]


def decode_y(sec, b_f32, c, h, w, win_u16=None):
    """b_f32: [C, h, w] float32 Laplace scales (host), win_u16: optional [C, h, w, 8] uint16 CDF windows
    from aivc_laplace_window -> q [C, h, w] int16."""
    L = _lib.lib()
    q = np.zeros((c, h, w), dtype=np.int16)
    idx, payload = y_channels(sec)
    if idx:
        all_ch = len(idx) == c                      # (the usual case: no gather of the per-symbol side data)
        scales = (b_f32 if all_ch else np.ascontiguousarray(b_f32[idx])).reshape(-1)
        sym = np.empty(scales.size, dtype=np.int16)
        buf = np.frombuffer(payload, dtype=np.uint8) if len(payload) else np.zeros(1, np.uint8)
        if win_u16 is not None:
            win = (win_u16 if all_ch else np.ascontiguousarray(win_u16[idx])).reshape(-1)
            _lib.check(L.aivc_rc_decode_laplace_win(scales.ctypes.data, win.ctypes.data, buf.ctypes.data,
                                                    len(payload), scales.size, sym.ctypes.data))
        else:
            _lib.check(L.aivc_rc_decode_laplace(scales.ctypes.data, buf.ctypes.data, len(payload),
                                                scales.size, sym.ctypes.data))
        q[idx] = sym.reshape(len(idx), h, w)
    return q
