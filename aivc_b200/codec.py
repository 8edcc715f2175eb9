"""Per-frame AIVC encoder / decoder on one B200.

Data flow of one inter frame (I frames skip the MOFNet half), all on one CUDA stream:

  planes (u8) --yuv420_to_fmap--> mof_in [code|prev|next] --g_a--> y --h_a--> z^ --h_s--> (mu, log var)
                                                                   \\--quantize_latent--> q, CDF bounds, y^
        g_a_ref(prev,next) --------------------------------> [y^ | shortcut] --g_s--> alpha,beta,v_prev,v_next
  warp_blend(prev, next) --> pred (-> codec_in [code|pred]), skip
  CodecNet: same chain on codec_in, shortcut = g_a_ref(pred)
  finalize_frame(codec_out + skip) --> 8-bit 4:2:0 planes (the reference for later frames)

It follows real_life/decode.py:455-898 (decoder) and its mirror for the encoder
(SURVEY.md 8a-19); entropy coding is handed to the host range coder (entropy.py) through
int16 symbols + 16-bit CDF bounds (encoder) or fp32 Laplace scales (decoder).
"""
import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib, entropy
from ._lib import F32
from .gop import FRAME_I, FRAME_P, FRAME_B, coding_order
from .plan import Plan, Buffer, Config


def _ceil_half(v, n):
    for _ in range(n):
        v = (v + 1) // 2
    return v


def latent_dims(h, w):
    return (_ceil_half(h, 4), _ceil_half(w, 4)), (_ceil_half(h, 6), _ceil_half(w, 6))


def _gain_of(net, frame_type):
    if not net.flag_gain_p_b or frame_type == FRAME_I:
        return net.gain_I
    return net.gain_P if frame_type == FRAME_P else net.gain_B


def _gain_vec(gm, idx_rate, mode, c):
    """abs(gain) of `mode` as `c` per-channel floats; a GainMatrix built with scalar_gain=True holds ONE value
    for all channels (gain_matrix.py:92-120) -- the kernels index the vector by channel, so it is expanded."""
    if hasattr(gm, 'gain_vector'):
        g = gm.gain_vector(idx_rate, mode).reshape(-1)
    else:
        g = gm.interpolate_gain_vector(idx_rate, mode=mode).detach().reshape(-1)
    if g.numel() == 1:
        g = g.expand(c)
    if g.numel() != c:
        raise ValueError('GainMatrix holds %d gains for a %d-channel latent' % (g.numel(), c))
    return g.float().contiguous()


class CondNetEngine:
    """Kernel plans and staging buffers of one ConditionalNet at a fixed frame size."""

    def __init__(self, net, h, w, in_buf, in_c, ref_off, ref_c, device, cfg, idx_rate=0., levels=False,
                 in_exact=False):
        self.net, self.device = net, device
        cy, cz, csc = net.nb_ft_y, net.nb_ft_z, net.out_c_shortcut_y
        self.cy, self.cz, self.csc = cy, cz, csc
        (hy, wy), (hz, wz) = latent_dims(h, w)
        self.dims_y, self.dims_z = (hy, wy), (hz, wz)
        exact = cfg.hyper_cfg()                    # hyperprior engine ('fp32' keeps sigma/mu exact)
        tc = cfg.tc
        embed = (lambda off: (in_buf.c, off, (1.0 / 255.0) if levels else 1.0)) if tc else (lambda off: None)
        self.g_s = Plan(net.g_s, hy, wy, cy + csc, device, cfg, pad_cout=16 if tc else 0)
        self.gs_in = self.g_s.src.buf
        self.has_ref = getattr(net, 'g_a_ref', None) is not None
        self.h_s = Plan(net.h_s, hz, wz, cz, device, exact)
        if getattr(net, 'g_a', None) is not None:        # a decoder-only model has no analysis side
            self.g_a = Plan(net.g_a, h, w, in_c, device, cfg, src_buf=in_buf,
                            out_scale=torch.ones(cy), in_embed=embed(0), in_exact=in_exact)
            # y stays fp32 for the quantiser; a tensor-core h_a reads a bf16, bordered copy of it
            self.h_a = Plan(net.h_a, hy, wy, cy, device, exact,
                            src_buf=self.g_a.dst.buf if exact.precision == 'fp32' else None,
                            dst_into=(self.h_s.src.buf, 0), out_post='round_clamp')
            self.y_copy = exact.precision != 'fp32' 
        if self.has_ref:
            self.g_a_ref = Plan(net.g_a_ref, h, w, ref_c, device, cfg, src_buf=in_buf,
                                src_c_off=0 if tc else ref_off, dst_into=(self.gs_in, cy),
                                in_embed=embed(ref_off), in_exact=in_exact)
        self.table = entropy.z_table_u16(net.pdf_z)
        self.gains = {}
        for ft in (FRAME_I, FRAME_P, FRAME_B):
            gm = _gain_of(net, ft)
            self.gains[ft] = (_gain_vec(gm, idx_rate, 'enc', cy).to(device).contiguous(),
                              _gain_vec(gm, idx_rate, 'dec', cy).to(device).contiguous())
        n_y, n_z = cy * hy * wy, cz * hz * wz
        self.n_y, self.n_z = n_y, n_z
        dev = dict(device=device)
        self.q_dev = torch.empty(n_y, dtype=torch.int16, **dev)
        self.bounds_dev = torch.empty(n_y, dtype=torch.int32, **dev)
        self.nz_dev = torch.zeros(cy, dtype=torch.int32, **dev)
        self.z_dev = torch.empty(n_z, dtype=torch.int16, **dev)
        self.b_dev = torch.empty(n_y, dtype=torch.float32, **dev)
        self.win_dev = torch.empty(n_y * 8, dtype=torch.int16, **dev)     # uint16 CDF windows (decoder)
        self._slots = []
        self._side = None
        self.overlap_shortcut = os.environ.get('AIVC_NO_OVERLAP') is None     # A/B switch

    # -- host staging: one pinned slot per latent in flight, so the GPU never waits for the coder
    def slot(self, i):
        while len(self._slots) <= i:
            pin = dict(pin_memory=True)
            self._slots.append(SimpleNamespace(
                z=torch.empty(self.n_z, dtype=torch.int16, **pin),
                # decoder: this latent's z on the device.  The hyper-decoder is re-run from it at synthesis
                # time; re-reading the pinned host copy there would race with the host rewriting the slot
                # for the next GOP while this GOP's reconstruction is still queued on the stream
                z_keep=torch.empty(self.n_z, dtype=torch.int16, device=self.device),
                bounds=torch.empty(self.n_y, dtype=torch.int32, **pin),
                nz=torch.empty(self.cy, dtype=torch.int32, **pin),
                b=torch.empty(self.n_y, dtype=torch.float32, **pin), win=None,
                q=torch.empty(self.n_y, dtype=torch.int16, **pin),
                event=torch.cuda.Event(), sec_y=None, first_of_i_frame=False))
        return self._slots[i]

    def _yhat_view(self):
        return self.gs_in.view(0, self.cy)

    def _shortcut(self, use):
        if use and self.has_ref:
            self.g_a_ref.run()
        else:   # zeros (decode.py:891-892), border included
            self.gs_in.zero_channels(self.cy, self.gs_in.c)

    # ---------------------------------------------------------------- encoder
    def encode_launch(self, sl, frame_type, use_shortcut, first_of_i_frame=False, rate=None):
        """Enqueue analysis, quantisation and synthesis; symbols / CDF bounds are copied to the
        pinned slot `sl` asynchronously.  No host synchronisation.  The shortcut transform g_a_ref does not
        depend on the analysis side, so it runs on a second stream next to g_a -> h_a -> h_s -> quantise and
        fills the SMs those leave idle (wave tails, the latency-bound 68x120 stages); g_s joins both.
        `rate`: optional fp32 device tensor [cy * hy * wy] receiving the per-symbol rate estimate in bits."""
        L, st = _lib.lib(), _lib.stream_ptr()
        enc_gain, dec_gain = self.gains[frame_type]
        main = torch.cuda.current_stream()
        overlap = self.overlap_shortcut and use_shortcut and self.has_ref
        if overlap:
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
                self._fork, self._join = torch.cuda.Event(), torch.cuda.Event()
            self._fork.record(main)
            self._side.wait_event(self._fork)
            with torch.cuda.stream(self._side):
                self._shortcut(True)
                self._join.record(self._side)
        self.g_a.ops[len(self.g_a.ops) - 1].out_scale = enc_gain.data_ptr()
        self.g_a.run()
        if self.y_copy:
            ys, yd = self.g_a.out_fmap, self.h_a.in_fmap
            _lib.check(L.aivc_fmap_copy(C.byref(ys), C.byref(yd), st))
        self.h_a.run()
        zf = self.h_s.in_fmap
        _lib.check(L.aivc_fmap_to_i16(C.byref(zf), self.z_dev.data_ptr(), st))
        sl.z.copy_(self.z_dev, non_blocking=True)
        self.h_s.run()
        self.nz_dev.zero_()
        yf, hsf, yh = self.g_a.out_fmap, self.h_s.out_fmap, self._yhat_view()
        _lib.check(L.aivc_quantize_latent(C.byref(yf), C.byref(hsf), dec_gain.data_ptr(),
                                          self.q_dev.data_ptr(), self.bounds_dev.data_ptr(),
                                          self.nz_dev.data_ptr(), C.byref(yh),
                                          None if rate is None else rate.data_ptr(), st))
        sl.bounds.copy_(self.bounds_dev, non_blocking=True)
        sl.nz.copy_(self.nz_dev, non_blocking=True)
        sl.event.record()
        sl.first_of_i_frame = first_of_i_frame
        if overlap:
            main.wait_event(self._join)
        else:
            self._shortcut(use_shortcut)
        self.g_s.run()

    def encode_finish(self, sl, keep=None):
        """Host side (any thread): range-code the latent of slot `sl` -> two bitstream sections.
        `keep`: optional dict that receives a copy of the z symbols (rate logging)."""
        sl.event.synchronize()
        (hy, wy), (hz, wz) = self.dims_y, self.dims_z
        if keep is not None:
            keep['z'] = sl.z.numpy().reshape(self.cz, hz, wz).copy()
        sec_z = entropy.encode_z(self.table, sl.z.numpy().reshape(self.cz, hz, wz))
        sec_y = entropy.encode_y(sl.bounds.numpy().view(np.uint32).reshape(self.cy, hy * wy),
                                 sl.nz.numpy())
        if sl.first_of_i_frame:    # bitstream.py:292-296
            sec_z = (0).to_bytes(4, 'big') * 2 + sec_z
        return sec_z + sec_y

    def encode(self, frame_type, use_shortcut, first_of_i_frame=False):
        sl = self.slot(0)
        self.encode_launch(sl, frame_type, use_shortcut, first_of_i_frame)
        return self.encode_finish(sl)

    # ---------------------------------------------------------------- decoder
    def decode_z_host(self, sec_z):
        """Host side (any thread): range-decode the z section."""
        (hz, wz) = self.dims_z
        return entropy.decode_z(self.table, sec_z, self.cz, hz, wz)

    def entropy_launch(self, sl, sec_z, sec_y, z=None):
        """z: host range decode (or the already decoded `z`) -> device -> h_s -> Laplace scales back to the
        pinned slot."""
        L, st = _lib.lib(), _lib.stream_ptr()
        (hy, wy), (hz, wz) = self.dims_y, self.dims_z
        if z is None:
            z = entropy.decode_z(self.table, sec_z, self.cz, hz, wz)
        # the previous user of this pinned slot has been waited for (its event was synchronised by
        # entropy_finish / encode_finish before the GOP call returned), so the host may rewrite it
        sl.z.copy_(torch.from_numpy(z).reshape(-1))
        sl.z_keep.copy_(sl.z, non_blocking=True)
        sl.sec_y = sec_y
        self._hyper_from_slot(sl)
        hs_y = self._hs_view()
        if sl.win is None:                          # pinned, allocated on the first decode only
            sl.win = torch.empty(self.n_y * 8, dtype=torch.int16, pin_memory=True)
        _lib.check(L.aivc_laplace_window(C.byref(hs_y), self.cy, self.b_dev.data_ptr(), self.win_dev.data_ptr(), st))
        sl.b.copy_(self.b_dev, non_blocking=True)
        sl.win.copy_(self.win_dev, non_blocking=True)
        sl.event.record()

    def _hs_view(self):
        hs_y = _lib.FMap.from_buffer_copy(self.h_s.out_fmap)
        hs_y.h, hs_y.w = self.dims_y                  # h_s(z)[:, :, :h_y, :w_y]  decode.py:853
        return hs_y

    def _hyper_from_slot(self, sl):
        L, st = _lib.lib(), _lib.stream_ptr()
        zf = self.h_s.in_fmap
        _lib.check(L.aivc_i16_to_fmap(sl.z_keep.data_ptr(), C.byref(zf), st))
        self.h_s.run()

    def entropy_finish(self, sl):
        """Host side (any thread): range-decode y of slot `sl` into its pinned q buffer."""
        sl.event.synchronize()
        (hy, wy) = self.dims_y
        q = entropy.decode_y(sl.sec_y, sl.b.numpy().reshape(self.cy, hy, wy), self.cy, hy, wy,
                             sl.win.numpy().view(np.uint16).reshape(self.cy, hy, wy, 8))
        sl.q.copy_(torch.from_numpy(q).reshape(-1))
        return True

    def synth_launch(self, sl, frame_type, use_shortcut, rerun_hyper=True):
        """y^ = (q + mu) * gain, shortcut, g_s.  `rerun_hyper`: recompute h_s(z) of this slot
        (another latent may have used the hyper-decoder buffers since entropy_launch)."""
        L, st = _lib.lib(), _lib.stream_ptr()
        _, dec_gain = self.gains[frame_type]
        if rerun_hyper:
            self._hyper_from_slot(sl)
        self.q_dev.copy_(sl.q, non_blocking=True)
        hs_y, yh = self._hs_view(), self._yhat_view()
        _lib.check(L.aivc_dequantize_latent(self.q_dev.data_ptr(), C.byref(hs_y), dec_gain.data_ptr(),
                                            C.byref(yh), st))
        self._shortcut(use_shortcut)
        self.g_s.run()

    def decode(self, sec_z, sec_y, frame_type, use_shortcut):
        sl = self.slot(0)
        self.entropy_launch(sl, sec_z, sec_y)
        self.entropy_finish(sl)
        self.synth_launch(sl, frame_type, use_shortcut, rerun_hyper=False)


class FrameCodec:
    """Encoder + decoder of one model at one frame size on one device."""

    def __init__(self, model, h, w, device='cuda:0', cfg=None, idx_rate=0., decoder_only=False):
        self.h, self.w = h, w
        self.device = torch.device(device)
        self.cfg = cfg or Config()
        self.idx_rate = idx_rate
        self._model = model
        # tensor-core engines: pixel inputs live in 16-channel (split-)bf16 pixels (zero padded) holding 8-bit
        # LEVEL units -- exact in bf16 -- with a 2-pixel replicate border for the 5x5 first conv; the
        # 1/255 is folded into the first-layer weights.  fp32 engine: plain [0,1] fp32 pixels.
        self.levels = self.cfg.tc
        with torch.cuda.device(self.device):
            if self.levels:
                self.mof_in = Buffer(h, w, 16, 2, self.cfg.act_dtype, self.device)
                self.codec_in = Buffer(h, w, 16, 2, self.cfg.act_dtype, self.device)
            else:
                self.mof_in = Buffer(h, w, 9, 0, F32, self.device)
                self.codec_in = Buffer(h, w, 6, 0, F32, self.device)
            self.skip = Buffer(h, w, 3, 0, F32, self.device)
            # (MOFNet reads code | prev | next: 8-bit levels, exact in bf16 when the frames arrive as uint8 planes)
            self.mof = CondNetEngine(model.mode_net.mode_net, h, w, self.mof_in, 9, 3, 6, self.device,
                                     self.cfg, idx_rate, self.levels, in_exact=self.levels)
            self.codec = CondNetEngine(model.codec_net.codec_net, h, w, self.codec_in, 6, 3, 3,
                                       self.device, self.cfg, idx_rate, self.levels)
            hc, wc = (h + 1) // 2, (w + 1) // 2
            self.zero_planes = (torch.zeros(h * w, dtype=torch.uint8, device=self.device),
                                torch.zeros(hc * wc, dtype=torch.uint8, device=self.device),
                                torch.zeros(hc * wc, dtype=torch.uint8, device=self.device))

    # -- helpers
    def new_planes(self):
        hc, wc = (self.h + 1) // 2, (self.w + 1) // 2
        d = dict(dtype=torch.uint8, device=self.device)
        return (torch.empty(self.h * self.w, **d), torch.empty(hc * wc, **d), torch.empty(hc * wc, **d))

    def _pack(self, planes, buf, c_off):
        y, u, v = planes
        u8 = 1 if y.dtype == torch.uint8 else 0
        if not u8 and buf is self.mof_in and self.levels:      # float planes need not be 8-bit levels: full split product
            for pl in (getattr(self.mof, 'g_a', None), getattr(self.mof, 'g_a_ref', None)):
                if pl is not None:
                    pl.set_input_exact(False)
        fm = buf.view(c_off, 3)
        _lib.check(_lib.lib().aivc_yuv420_to_fmap(y.data_ptr(), u.data_ptr(), v.data_ptr(), u8,
                                                  1 if self.levels else 0, C.byref(fm),
                                                  _lib.stream_ptr()))

    def _pack16(self, code, prev, nxt, into_codec):
        """tensor-core engines, uint8 planes: [code | prev | next] -> mof_in (and code -> codec_in) in ONE launch.
        None = the all-zero frame."""
        ptrs = []
        for planes in (code, prev, nxt):
            ptrs += [None, None, None] if planes is None else [p.data_ptr() for p in planes]
        d1 = self.mof_in.view(0, 16)
        d2 = self.codec_in.view(0, 16) if into_codec else None
        _lib.check(_lib.lib().aivc_yuv420_pack16(*ptrs, C.byref(d1), None if d2 is None else C.byref(d2),
                                                 _lib.stream_ptr()))

    def _fusable(self, *frames):
        return self.levels and all(f is None or f[0].dtype == torch.uint8 for f in frames)

    def _zero_pred(self):
        self.codec_in.zero_channels(3, 6)

    def _motion(self, frame_type, aux=None):
        """aux: optional fp32 device tensor [5, h, w] receiving alpha, beta and x_warp (3 planes, [0,1] units)."""
        L = _lib.lib()
        mo = self.mof.g_s.out_fmap
        prev, nxt = self.mof_in.view(3, 3), self.mof_in.view(6, 3)
        pred, skip = self.codec_in.view(3, 3), self.skip.view(0, 3)
        _lib.check(L.aivc_warp_blend(C.byref(mo), C.byref(prev), C.byref(nxt),
                                     1 if frame_type == FRAME_P else 0, 1 if self.levels else 0,
                                     C.byref(pred), C.byref(skip), None if aux is None else aux.data_ptr(),
                                     _lib.stream_ptr()))

    def _finalize(self, frame_type, out_planes):
        cod = _lib.FMap.from_buffer_copy(self.codec.g_s.out_fmap)
        cod.h, cod.w = self.h, self.w                  # crop (decode.py:652)
        sk = self.skip.view(0, 3)
        y, u, v = out_planes
        _lib.check(_lib.lib().aivc_finalize_frame(
            C.byref(cod), None if frame_type == FRAME_I else C.byref(sk), y.data_ptr(), u.data_ptr(),
            v.data_ptr(), None, _lib.stream_ptr()))

    def _refs(self, frame_type, prev_rec, next_rec):
        self._pack(prev_rec if frame_type != FRAME_I else self.zero_planes, self.mof_in, 3)
        self._pack(next_rec if frame_type == FRAME_B else self.zero_planes, self.mof_in, 6)

    # -- public
    def encode_frame(self, planes, frame_type, prev_rec=None, next_rec=None):
        """planes: (y, u, v) flat device tensors, uint8 levels or fp32 in [0,1].
        Returns (frame bitstream bytes, reconstructed uint8 planes)."""
        with torch.cuda.device(self.device):
            data = b''
            self._pack(planes, self.codec_in, 0)
            if frame_type == FRAME_I:
                self._zero_pred()
            else:
                self._pack(planes, self.mof_in, 0)
                self._refs(frame_type, prev_rec, next_rec)
                data += self.mof.encode(frame_type, frame_type == FRAME_B)
                self._motion(frame_type)
            data += self.codec.encode(frame_type, frame_type != FRAME_I,
                                      first_of_i_frame=(frame_type == FRAME_I))
            rec = self.new_planes()
            self._finalize(frame_type, rec)
        return data, rec

    def decode_frame(self, frame_bytes, frame_type, prev_rec=None, next_rec=None):
        """Decoder.decode (decode.py:455-580). Returns reconstructed uint8 planes."""
        with torch.cuda.device(self.device):
            secs = entropy.split_sections(frame_bytes)
            if frame_type == FRAME_I:
                self._zero_pred()
            else:
                self._refs(frame_type, prev_rec, next_rec)
                self.mof.decode(secs[0], secs[1], frame_type, frame_type == FRAME_B)
                self._motion(frame_type)
            self.codec.decode(secs[2], secs[3], frame_type, frame_type != FRAME_I)
            rec = self.new_planes()
            self._finalize(frame_type, rec)
        return rec

    # ------------------------------------------------------------------ GOP level: frames in flight
    # Frames of one dependency level of a GOP (gop.levels: e.g. 1, 1, 1, 2, 4, 8, 16 frames for '1_GOP_32') are
    # independent given the reconstructions of earlier levels.  With cfg.frames_in_flight = 2 the frames of a level
    # are dealt to two LANES -- this codec on the caller's stream and a second set of plans / buffers on its own
    # stream -- so the GPU always has two independent kernel sequences to draw from: the wave tail of one frame's
    # layer (3.65 waves of tiles on 148 SMs at 270x480) is filled by the other frame's CTAs, and a frame waiting for
    # its symbols from the host range decoder does not idle the device.  Every frame is still coded by the same
    # kernels in the same order, so bitstream and reconstruction do not depend on the number of lanes.
    def _lanes(self):
        n = max(1, int(self.cfg.frames_in_flight)) if getattr(self, 'lanes_enabled', True) else 1
        if n > 1 and getattr(self, '_lane_codecs', None) is None:
            import dataclasses
            solo = dataclasses.replace(self.cfg, frames_in_flight=1)
            self._lane_codecs = [FrameCodec(self._model, self.h, self.w, self.device, solo, self.idx_rate)
                                 for _ in range(n - 1)]
            self._lane_streams = [torch.cuda.Stream(device=self.device) for _ in range(n - 1)]
        return [self] + (self._lane_codecs if n > 1 else [])

    def _schedule(self, gop):
        """[(frame, lane)] in an order that respects dependencies: level by level, frames of a level dealt
        round-robin to the lanes (a level of one frame stays on lane 0)."""
        from .gop import levels
        n = len(self._lanes())
        out = []
        for level in levels(gop):
            for i, f in enumerate(level):
                out.append((f, i % n if len(level) > 1 else 0))
        return out

    def _lane_ctx(self, lane, start_ev):
        """stream context of a lane; lane > 0 streams first wait for everything queued on the caller's stream"""
        if lane == 0:
            import contextlib
            return contextlib.nullcontext()
        st = self._lane_streams[lane - 1]
        if start_ev is not None and not start_ev[1][lane]:
            st.wait_event(start_ev[0])
            start_ev[1][lane] = True
        return torch.cuda.stream(st)

    def _wait_refs(self, lane, done, owner, *refs):
        """the current stream (of `lane`) waits for reference frames reconstructed on another lane"""
        for r in refs:
            if r is not None and owner.get(r, lane) != lane:
                torch.cuda.current_stream().wait_event(done[r])

    def _launch_encode_frame(self, slot_i, f, frames, gop, rec, pool, aux):
        """Enqueue one frame of a GOP on the current stream -> (futures of its bitstream sections, planes)."""
        e = gop[f]
        ft = e['type']
        parts = []
        f32 = dict(dtype=torch.float32, device=self.device)
        prev_r = rec.get(e['prev_ref']) if ft != FRAME_I else None
        next_r = rec.get(e['next_ref']) if ft == FRAME_B else None
        fused = ft != FRAME_I and self._fusable(frames[f], prev_r, next_r)
        if fused:        # [code | prev | next] and the CodecNet's code in one launch
            self._pack16(frames[f], prev_r, next_r, True)
        else:
            self._pack(frames[f], self.codec_in, 0)
        if ft == FRAME_I:
            self._zero_pred()
        elif not fused:
            self._pack(frames[f], self.mof_in, 0)
            self._refs(ft, prev_r, next_r)
        a = None
        if aux is not None:
            a = aux[f] = {'mode_keep': {}, 'codec_keep': {}, 'codec_rate_y': torch.empty(self.codec.n_y, **f32)}
            if ft != FRAME_I:
                a['mode_rate_y'] = torch.empty(self.mof.n_y, **f32)
                a['warp'] = torch.empty(5 * self.h * self.w, **f32)
        if ft != FRAME_I:
            sl = self.mof.slot(slot_i)
            self.mof.encode_launch(sl, ft, ft == FRAME_B, rate=a and a['mode_rate_y'])
            parts.append(pool.submit(self.mof.encode_finish, sl, a and a['mode_keep']))
            self._motion(ft, a and a['warp'])
        sl = self.codec.slot(slot_i)
        self.codec.encode_launch(sl, ft, ft != FRAME_I, first_of_i_frame=(ft == FRAME_I),
                                 rate=a and a['codec_rate_y'])
        parts.append(pool.submit(self.codec.encode_finish, sl, a and a['codec_keep']))
        planes = self.new_planes()
        self._finalize(ft, planes)
        return parts, planes

    def encode_gop(self, frames, gop, aux=None):
        """frames: {'frame_i': (y,u,v) device planes}. Returns ({name: bytes}, {name: planes}).
        The GPU runs ahead through the whole GOP; the serial range coder of every latent runs on
        host worker threads as soon as its symbols have landed in pinned memory.
        aux: optional dict; filled per frame with what the reference's encoder logs (loss_function.py:158-204):
        'mode_rate_y' / 'codec_rate_y' (fp32 [cy * hy * wy] device tensors, bits per symbol), 'mode_keep' / 'codec_keep'
        ({'z': int16 numpy symbols}, for the z rate) and 'warp' (fp32 [5 * h * w]: alpha, beta, x_warp; inter frames)."""
        pool = self._pool()
        lanes = self._lanes()
        rec, futs, done, owner = {}, {}, {}, {}
        count = [0] * len(lanes)
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream()
            start = (torch.cuda.Event(), [True] + [False] * (len(lanes) - 1))
            start[0].record(main)
            for f, lane in self._schedule(gop):
                e = gop[f]
                with self._lane_ctx(lane, start):
                    self._wait_refs(lane, done, owner, e['prev_ref'], e['next_ref'])
                    futs[f], rec[f] = lanes[lane]._launch_encode_frame(count[lane], f, frames, gop, rec, pool, aux)
                    if len(lanes) > 1:
                        done[f] = torch.cuda.Event()
                        done[f].record()
                        owner[f] = lane
                        for p in rec[f]:            # planes allocated under this lane's stream, read by all of them
                            p.record_stream(main)
                            for st in self._lane_streams:
                                p.record_stream(st)
                count[lane] += 1
            for lane in range(1, len(lanes)):       # the caller's stream sees everything the other lanes produced
                if start[1][lane]:
                    ev = torch.cuda.Event()
                    ev.record(self._lane_streams[lane - 1])
                    main.wait_event(ev)
        out_b = {f: b''.join(p.result() for p in futs[f]) for f in coding_order(gop)}
        return out_b, rec

    def decode_gop(self, frame_bytes, gop, lookahead=None):
        """Entropy decoding runs ahead of reconstruction: a latent's z is range-decoded by a host worker
        (they depend on nothing and all start at once), its hyper-decoder runs on the GPU, the Laplace
        scales + CDF windows go back to pinned memory and a worker decodes y -- none of this depends on
        reconstructed pixels.  Reconstruction follows level by level (see `_lanes`), consuming the symbols as the
        workers deliver them.  `lookahead` is kept for API compatibility (the whole GOP's entropy stages are
        always enqueued first: measured 257 ms per 1080p GOP against 270 ms with 8 frames interleaved)."""
        pool = self._pool()
        lanes = self._lanes()
        sched = self._schedule(gop)
        futs, zf, secs_of = {}, {}, {}
        for f, _ in sched:
            secs_of[f] = entropy.split_sections(frame_bytes[f])
            if gop[f]['type'] != FRAME_I:
                zf[(f, 0)] = pool.submit(self.mof.decode_z_host, secs_of[f][0])
            zf[(f, 1)] = pool.submit(self.codec.decode_z_host, secs_of[f][2])
        rec, done, owner, slot_of = {}, {}, {}, {}
        count = [0] * len(lanes)
        self.last_slots = {}
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream()
            start = (torch.cuda.Event(), [True] + [False] * (len(lanes) - 1))
            start[0].record(main)
            for f, lane in sched:                    # pass 1: hyper-decoders + CDF windows, y decoding on workers
                c = lanes[lane]
                i = slot_of[f] = count[lane]
                count[lane] += 1
                secs = secs_of[f]
                with self._lane_ctx(lane, start):
                    if gop[f]['type'] != FRAME_I:
                        sl = c.mof.slot(i)
                        c.mof.entropy_launch(sl, secs[0], secs[1], z=zf[(f, 0)].result())
                        futs[(f, 0)] = pool.submit(c.mof.entropy_finish, sl)
                        self.last_slots[(f, 'mof')] = sl
                    sl = c.codec.slot(i)
                    c.codec.entropy_launch(sl, secs[2], secs[3], z=zf[(f, 1)].result())
                    futs[(f, 1)] = pool.submit(c.codec.entropy_finish, sl)
                    self.last_slots[(f, 'codec')] = sl
            for f, lane in sched:                    # pass 2: reconstruction
                c, i = lanes[lane], slot_of[f]
                e = gop[f]
                ft = e['type']
                with self._lane_ctx(lane, start):
                    self._wait_refs(lane, done, owner, e['prev_ref'], e['next_ref'])
                    if ft == FRAME_I:
                        c._zero_pred()
                    else:
                        prev_r, next_r = rec.get(e['prev_ref']), rec.get(e['next_ref']) if ft == FRAME_B else None
                        if c._fusable(prev_r, next_r):
                            c._pack16(None, prev_r, next_r, False)
                        else:
                            c._refs(ft, prev_r, next_r)
                        futs[(f, 0)].result()
                        c.mof.synth_launch(c.mof.slot(i), ft, ft == FRAME_B)
                        c._motion(ft)
                    futs[(f, 1)].result()
                    c.codec.synth_launch(c.codec.slot(i), ft, ft != FRAME_I)
                    rec[f] = c.new_planes()
                    c._finalize(ft, rec[f])
                    if len(lanes) > 1:
                        done[f] = torch.cuda.Event()
                        done[f].record()
                        owner[f] = lane
                        for p in rec[f]:
                            p.record_stream(main)
                            for st in self._lane_streams:
                                p.record_stream(st)
            for lane in range(1, len(lanes)):
                if start[1][lane]:
                    ev = torch.cuda.Event()
                    ev.record(self._lane_streams[lane - 1])
                    main.wait_event(ev)
        return rec

    def decoded_symbols(self, f, net_name):
        """(q [cy, hy, wy], z [cz, hz, wz]) int16 numpy: the symbols the last decode_gop read for frame `f`
        ('mof' | 'codec'); parity tools compare them with the oracle's indices."""
        sl = self.last_slots[(f, net_name)]
        eng = self.mof if net_name == 'mof' else self.codec
        return (sl.q.numpy().reshape(eng.cy, *eng.dims_y).copy(), sl.z.numpy().reshape(eng.cz, *eng.dims_z).copy())

    def _pool(self):
        if getattr(self, '_tp', None) is None:
            # one process per GPU: the ranks of a node share its cores for the range coding
            per_node = max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1')))
            self._tp = ThreadPoolExecutor(max_workers=max(2, min(32, (os.cpu_count() or 4) // per_node)))
        return self._tp


def planes_to_device(yuv_u8, device):
    """(y, u, v) numpy uint8 arrays -> flat device tensors."""
    return tuple(torch.from_numpy(np.ascontiguousarray(p).reshape(-1)).to(device) for p in yuv_u8)


def gop_forward(model, model_input):
    from .adapter import gop_forward as _gf
    return _gf(model, model_input)
