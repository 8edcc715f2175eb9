"""Direct planar YUV 4:2:0 file path.

The reference goes .yuv -> three 8-bit PNGs per frame (format_conversion/yuv_to_png.py, one bash + dd +
python fork per frame) -> load_frames -> /255 floats, and back through PNGs to a .yuv
(src/encode.py:75, src/decode.py:100).  PNG is lossless, so the codec sees exactly the uint8 planes of the
raw file; here they are read straight from the file into pinned host memory and copied to the device
asynchronously, and decoded planes are appended straight to the output file.

  file name  : <name>_<W>x<H>_<fps>_420.yuv         format_conversion/utils.py:45-50, 69-72
  frame      : W*H luma bytes, then (W/2)*(H/2) U bytes, then V bytes; 8 bit         utils.py:95-103
  GOP split  : frames [start, end] in ceil(n / gop_size) GOPs, the last one padded by repeating the last
               frame (model_management.py:142-173, img_processing.py:155-172); padded frames are coded
               but dropped again by the decoder side through idx_first / idx_last of the video header.
"""
import os

import numpy as np
import torch

from . import container
from .gop import generate_gop_struct


def parse_name(path):
    """(width, height, fps) from '<name>_<W>x<H>_<fps>_420.yuv'."""
    toks = os.path.basename(path).split('_')
    try:
        w, h = (int(float(t)) for t in toks[1].split('x'))
        fps = float(toks[2])
    except (IndexError, ValueError):
        raise ValueError('cannot read the resolution from %r: expected <name>_<W>x<H>_<fps>_420.yuv' % path)
    return w, h, fps


def frame_bytes(w, h):
    return w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)


class YuvReader:
    """Random access to the frames of a planar 8-bit 4:2:0 file; `pinned(i)` stages a frame in a ring of
    pinned host buffers for an asynchronous host-to-device copy."""

    def __init__(self, path, w=None, h=None, ring=4):
        if w is None or h is None:
            w, h, self.fps = parse_name(path)
        else:
            self.fps = None
        self.path, self.w, self.h = path, w, h
        self.wc, self.hc = (w + 1) // 2, (h + 1) // 2
        self.fb = frame_bytes(w, h)
        size = os.path.getsize(path)
        if size % self.fb:
            raise ValueError('%s: %d bytes is not a whole number of %dx%d 4:2:0 frames' % (path, size, w, h))
        self.nb_frames = size // self.fb
        self._mm = np.memmap(path, dtype=np.uint8, mode='r') if size else np.zeros(0, np.uint8)
        self._ring, self._slot = [None] * ring, 0

    def __len__(self):
        return self.nb_frames

    def frame(self, i):
        """(y, u, v) flat uint8 numpy views of frame i (no copy)."""
        if not 0 <= i < self.nb_frames:
            raise IndexError('frame %d of %d' % (i, self.nb_frames))
        base, ny, nc = i * self.fb, self.w * self.h, self.wc * self.hc
        return (self._mm[base:base + ny], self._mm[base + ny:base + ny + nc],
                self._mm[base + ny + nc:base + ny + 2 * nc])

    def pinned(self, i):
        """Frame i copied into the next pinned ring slot -> (y, u, v) flat uint8 torch tensors."""
        k = self._slot
        self._slot = (k + 1) % len(self._ring)
        if self._ring[k] is None:
            pin = torch.cuda.is_available()
            self._ring[k] = torch.empty(self.fb, dtype=torch.uint8, pin_memory=pin)
        buf = self._ring[k]
        buf.numpy()[:] = self._mm[i * self.fb:(i + 1) * self.fb]
        ny, nc = self.w * self.h, self.wc * self.hc
        return buf[:ny], buf[ny:ny + nc], buf[ny + nc:ny + 2 * nc]

    def gop_frames(self, first, gop_size, last_valid, device):
        """{'frame_k': device planes} for the GOP starting at absolute frame `first`; frames past
        `last_valid` repeat it (the reference's padding of an incomplete last GOP)."""
        # one pinned slot per frame of the GOP: a slot is only rewritten for the NEXT GOP, after encode_gop
        # has returned (it waits for every frame's symbols, which are stream-ordered behind these copies)
        if len(self._ring) < gop_size:
            self._ring += [None] * (gop_size - len(self._ring))
        self._slot = 0
        out = {}
        for k in range(gop_size):
            i = min(first + k, last_valid)
            out['frame_%d' % k] = tuple(p.to(device, non_blocking=True) for p in self.pinned(i))
        return out


class YuvWriter:
    def __init__(self, path):
        self.f = open(path, 'wb')
        self.nb_frames = 0

    def append(self, planes):
        """planes: (y, u, v) flat uint8 tensors (any device) or numpy arrays."""
        for p in planes:
            a = p.detach().cpu().numpy() if isinstance(p, torch.Tensor) else np.asarray(p)
            self.f.write(np.ascontiguousarray(a, dtype=np.uint8).tobytes())
        self.nb_frames += 1

    def close(self):
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def gop_schedule(idx_start, idx_end, gop_size):
    """[(first absolute frame, frames to keep)] per GOP -- model_management.py:142-173."""
    n = idx_end - idx_start + 1
    nb_gop = -(-n // gop_size)
    return [(idx_start + g * gop_size, min(gop_size, n - g * gop_size)) for g in range(nb_gop)]


def encode_yuv_file(codec_for, model, path_in, gop_name, idx_start=0, idx_end=-1, device='cuda:0', cfg=None,
                    idx_rate=0.):
    """.yuv file -> AIVC video bitstream bytes (container.pack_video), GOP by GOP.
    `codec_for(model, h, w, device, cfg, idx_rate)` supplies the FrameCodec (adapter.codec_for)."""
    from .codec import latent_dims
    rd = YuvReader(path_in)
    if idx_end < 0:
        idx_end = len(rd) - 1
    if not 0 <= idx_start <= idx_end < len(rd):
        raise ValueError('frames [%d, %d] outside the %d frames of %s' % (idx_start, idx_end, len(rd), path_in))
    gop = generate_gop_struct(gop_name)
    order = sorted(gop, key=lambda f: int(f.split('_')[1]))
    dev = torch.device(device)
    codec = codec_for(model, rd.h, rd.w, dev, cfg, idx_rate)
    gops = []
    for first, _keep in gop_schedule(idx_start, idx_end, len(order)):
        frames = rd.gop_frames(first, len(order), idx_end, dev)
        bts, _rec = codec.encode_gop(frames, gop)
        gops.append(container.pack_gop(gop_name, [bts[f] for f in order], idx_rate))
    dim_y, dim_z = latent_dims(rd.h, rd.w)
    return container.pack_video((rd.h, rd.w), dim_y, dim_z, gops, idx_start, idx_end)


def decode_to_yuv_file(decode_video, model, video_bytes, path_out, device='cuda:0', cfg=None):
    """AIVC video bitstream -> planar .yuv; the padding frames of an incomplete last GOP are dropped
    (idx_first / idx_last of the video header).  `decode_video` = adapter.decode_video."""
    gops, _dims, first, last = decode_video(model, video_bytes, device, cfg)
    todo = last - first + 1
    with YuvWriter(path_out) as wr:
        for g in gops:
            for f in sorted(g, key=lambda n: int(n.split('_')[1])):
                if todo <= 0:
                    break
                wr.append(g[f])
                todo -= 1
        return wr.nb_frames
