"""Lowering of AIVC layer trees to fused convolution stages, and their execution.

A transform (``g_a``, ``g_s``, ``h_a``, ``h_s``, ``g_a_ref`` or any single layer) is a tree of
``Sequential`` / ``CustomConvLayer`` / ``UpscalingLayer`` / ``ChengResBlock`` / ``ResBlock`` /
``AttentionResBlock`` / ``SimplifiedAttention`` / ``Conv2d`` ... nodes.  ``lower`` walks it by
class *name* (so reference modules, un-pickled models and the mirrors in
``aivc_b200.layers`` are all accepted) and emits one ``aivc_conv_op`` per convolution with
everything that follows it folded into the epilogue:

    out = post( act(conv(in) + bias) * gate + residual ) * out_scale

  ChengResBlock plain/down/up  -> residual of the 2nd conv   (custom_conv_layers.py:105-109)
  ResBlock / AttentionResBlock -> residual + post ReLU/Leaky  (custom_conv_layers.py:126, attention.py:41)
  SimplifiedAttention          -> sigmoid act, gate = trunk, residual = x   (attention.py:90-97)
  GDN / IGDN                   -> act of the producing conv   (misc_layers.py:113-154)

Activations live in NHWC buffers with a replicate border (see include/aivc_b200.h); buffers are
recycled by liveness.  All stages of a plan run with ONE ctypes call
(``aivc_conv2d_fused_seq``) on the current CUDA stream.
"""
import ctypes as C
import weakref
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import FMap, ConvOp, ACT, POST, F32, BF16, F16, BF16X2, ENGINE_SIMT, ENGINE_TC, ENGINE_TC_X3

_TORCH_DT = {F32: torch.float32, BF16: torch.bfloat16, F16: torch.float16, BF16X2: torch.bfloat16}


@dataclass
class Config:
    precision: str = 'bf16x3'      # 'fp32'  : every stage on the exact SIMT engine, fp32 buffers
                                   # 'bf16x3': tcgen05 engine on split-bf16 operands (hi.Whi + lo.Whi + hi.Wlo per
                                   #           K chunk, fp32 accumulation): fp32-grade results, 3x the MMA work --
                                   #           the mode whose latent indices match the fp32 reference
                                   # 'bf16'  : tcgen05 engine on plain bf16 operands: fastest, ~1 % of the latent
                                   #           indices differ from fp32 arithmetic
    tc_min_cin: int = 16           # tensor-core stages need cin % 16 == 0 and cout % 16 == 0
    hyper: str = 'auto'            # precision of h_a / h_s: 'fp32' (exact engine), 'bf16', or
                                   # 'auto' = same as `precision`
    two_lanes: bool = True         # independent branches of attention blocks on two CUDA streams
    s2d_first: bool = True         # tensor-core engines: 5x5 stride-2 pixel-domain first layer as space-to-depth + 3x3 conv
    cuda_graphs: bool = False      # replay every transform as one CUDA graph (plan.Plan.run).  Measured on B200: no gain
                                   # (24.4 / 60.1 frames/s either way: the device is kernel-bound, the host keeps the
                                   # queue full), so plain launches stay the default
    frames_in_flight: int = 1      # GOP coding: independent frames of a dependency level on this many streams
                                   # (codec.py::_lanes).  Measured on B200 (tools/lanes_ab.py): 2 lanes give +1 % (bf16x3)
                                   # / +2 % (bf16) for twice the buffers -- the device is already 95 % busy and
                                   # power-capped with one frame in flight -- so one lane is the default

    def key(self):
        return (self.precision, self.tc_min_cin, self.hyper, self.two_lanes, self.s2d_first, self.frames_in_flight,
                self.cuda_graphs)

    @property
    def tc(self):
        return self.precision in ('bf16', 'bf16x3')

    @property
    def act_dtype(self):
        """dtype of tensors a tensor-core stage reads"""
        return BF16X2 if self.precision == 'bf16x3' else BF16

    def hyper_cfg(self):
        h = self.precision if self.hyper == 'auto' else self.hyper
        return Config(precision=h, tc_min_cin=self.tc_min_cin, hyper=h, two_lanes=self.two_lanes,
                      s2d_first=self.s2d_first, frames_in_flight=self.frames_in_flight, cuda_graphs=self.cuda_graphs)


DEFAULT = Config()


class Buffer:
    """One device allocation holding a bordered NHWC tensor."""

    def __init__(self, h, w, c, pad, dtype, device):
        self.h, self.w, self.c, self.pad, self.dtype = h, w, c, pad, dtype
        self.pitch = w + 2 * pad + ((w + 2 * pad) & 1)      # even: stride-2 TMA views pair columns
        self.rows = h + 2 * pad + ((h + 2 * pad) & 1)
        # split bf16: a pixel is [hi(c) | lo(c)], 2c elements
        self.c_stride = 2 * c if dtype == BF16X2 else c
        self.t = torch.zeros(self.rows * self.pitch * self.c_stride, dtype=_TORCH_DT[dtype], device=device)

    def view(self, c_off=0, c=None, h=None, w=None):
        return FMap(self.t.data_ptr(), self.h if h is None else h, self.w if w is None else w,
                    self.c - c_off if c is None else c, c_off, self.c_stride, self.pad, self.pitch,
                    self.rows, self.dtype, 0)

    def interior(self):
        """[h, w, c] torch view of the logical tensor (debug / tests; split bf16: the hi halves)."""
        full = self.t.view(self.rows, self.pitch, self.c_stride)
        return full[self.pad:self.pad + self.h, self.pad:self.pad + self.w, :self.c]

    def zero_channels(self, c0, c1):
        """channels [c0, c1) of every pixel, border included, := 0 (stream-ordered)"""
        full = self.t.view(self.rows, self.pitch, self.c_stride)
        full[:, :, c0:c1].zero_()
        if self.dtype == BF16X2:
            full[:, :, self.c + c0:self.c + c1].zero_()


@dataclass
class T:
    """Symbolic tensor of a plan."""
    h: int
    w: int
    c: int
    pad: int = 0
    dtype: int = F32
    buf: Optional[Buffer] = None
    c_off: int = 0
    external: bool = False
    first: int = 10 ** 9
    last: int = -1

    def fmap(self):
        return self.buf.view(self.c_off, self.c, self.h, self.w)


@dataclass
class Stage:
    kind: int
    k: int
    stride: int
    src: T
    dst: T
    weight: torch.Tensor            # original layout, fp32
    bias: Optional[torch.Tensor]
    act: str = 'no'
    post: str = 'none'
    gdn: Optional[tuple] = None     # (beta, gamma, inverse)
    res: Optional[T] = None
    gate: Optional[T] = None
    out_scale: Optional[torch.Tensor] = None
    engine: int = ENGINE_SIMT
    cin_off: int = 0                # weight input channel 0 sits at this channel of the source pixel
    w_scale: float = 1.0            # folded into the packed weights (1/255 for level-unit inputs)
    flags: int = 0                  # _lib.OP_LANE1 / OP_FORK / OP_JOIN
    alg_flops: float = 0.0          # FLOPs of the reference op(s) when the stage was rewritten (0: from geometry)


class Graph:
    def __init__(self):
        self.stages = []

    def conv(self, kind, k, stride, src, weight, bias, act='no', gdn=None):
        if kind == 0:
            cout = weight.shape[0]
            h, w = -(-src.h // stride), -(-src.w // stride)
        else:
            cout = weight.shape[1]
            h, w = 2 * src.h, 2 * src.w
        dst = T(h, w, cout)
        self.stages.append(Stage(kind, k, stride, src, dst, weight.detach(), None if bias is None
                                 else bias.detach(), act, gdn=gdn))
        return dst


def _act_of(m):
    n = type(m).__name__
    if n == 'LeakyReLU':
        if abs(m.negative_slope - 0.01) > 1e-12:
            raise NotImplementedError('LeakyReLU slope != 0.01')
        return 'leaky_relu'
    return {'ReLU': 'relu', 'Sigmoid': 'sigmoid'}.get(n)


def _gdn_params(m):
    dev = m.beta.device
    ped = m.pedestal.to(dev)
    beta = torch.maximum(m.beta.detach(), m.beta_bound.to(dev)) ** 2 - ped
    gamma = torch.maximum(m.gamma.detach(), m.gamma_bound.to(dev)) ** 2 - ped
    return beta.float(), gamma.float(), bool(m.inverse)


def _lower_seq(seq, x, g):
    children = list(seq)
    i = 0
    while i < len(children):
        m = children[i]
        n = type(m).__name__
        if n == 'ReplicationPad2d':
            nxt = children[i + 1]
            if type(nxt).__name__ != 'Conv2d' or m.padding[0] != nxt.kernel_size[0] // 2:
                raise NotImplementedError('ReplicationPad2d must feed a same-size Conv2d')
            i += 1
            continue
        if n in ('Conv2d', 'ConvTranspose2d'):
            k = m.kernel_size[0]
            if n == 'Conv2d':
                padded = i > 0 and type(children[i - 1]).__name__ == 'ReplicationPad2d'
                if k > 1 and not padded:
                    raise NotImplementedError('un-padded k>1 Conv2d does not occur in AIVC')
                kind, stride = 0, m.stride[0]
            else:
                if m.stride[0] != 2 or m.output_padding[0] != 1 or m.padding[0] != (k + 1) // 2 - 1:
                    raise NotImplementedError('ConvTranspose2d geometry other than AIVC\'s x2')
                kind, stride = 1, 2
            act, gdn = 'no', None
            if i + 1 < len(children):
                nxt = children[i + 1]
                if type(nxt).__name__ == 'GDN':
                    gdn = _gdn_params(nxt)
                    act = 'gdn_inverse' if gdn[2] else 'gdn'
                    i += 1
                elif _act_of(nxt) is not None:
                    act = _act_of(nxt)
                    i += 1
            x = g.conv(kind, k, stride, x, m.weight, m.bias, act, gdn)
        else:
            x = lower(m, x, g)
        i += 1
    return x


def s2d_weights(w16):
    """[cout, c, 5, 5] stride-2 kernel -> [cout, 4c, 3, 3] stride-1 kernel over the space-to-depth image
    (channel (dy*2+dx)*c + ch of block (by, bx) = pixel (2by+dy, 2bx+dx)): tap (ky, kx) = (2ty+dy, 2tx+dx);
    the 11 positions of the 6x6 footprint outside the 5x5 stay zero."""
    cout, c = w16.shape[0], w16.shape[1]
    w3 = torch.zeros(cout, 4 * c, 3, 3, dtype=torch.float32)
    for ty in range(3):
        for dy in range(2):
            if 2 * ty + dy > 4:
                continue
            for tx in range(3):
                for dx in range(2):
                    if 2 * tx + dx > 4:
                        continue
                    blk = (dy * 2 + dx) * c
                    w3[:, blk:blk + c, ty, tx] = w16[:, :, 2 * ty + dy, 2 * tx + dx]
    return w3


def _fold(g, res=None, gate=None, post='none'):
    s = g.stages[-1]
    assert s.res is None and s.gate is None and s.post == 'none', 'epilogue already occupied'
    s.res, s.gate, s.post = res, gate, post


def lower(m, x, g):
    n = type(m).__name__
    if n == 'Sequential':
        return _lower_seq(m, x, g)
    if n in ('CustomConvLayer', 'UpscalingLayer'):
        return _lower_seq(m.layers, x, g)
    if n in ('Conv2d', 'ConvTranspose2d'):
        return _lower_seq([m], x, g)
    if n == 'ChengResBlock':
        if m.mode == 'plain':
            skip = x
        else:
            skip = lower(m.aux_layer, x, g)
        y = _lower_seq(m.layers, x, g)
        _fold(g, res=skip)
        return y
    if n == 'ResBlock':
        y = _lower_seq(m.layers, x, g)
        _fold(g, res=x, post='relu')
        return y
    if n == 'AttentionResBlock':
        y = _lower_seq(m.layers, x, g)
        _fold(g, res=x, post='leaky_relu')
        return y
    if n == 'SimplifiedAttention':
        # trunk and attention paths are independent until the final gate: trunk runs on the side
        # lane (second CUDA stream), the attention path on the caller's stream
        i0 = len(g.stages)
        trunk = _lower_seq(m.trunk, x, g)
        for s in g.stages[i0:]:
            s.flags |= _lib.OP_LANE1
        g.stages[i0].flags |= _lib.OP_FORK
        y = _lower_seq(m.attention, x, g)
        _fold(g, res=x, gate=trunk)
        g.stages[-1].flags |= _lib.OP_JOIN
        return y
    if n == 'GDN':
        raise NotImplementedError('stand-alone GDN: wrap it with its producing conv')
    raise NotImplementedError('aivc_b200.plan: cannot lower ' + n)


class Plan:
    """Executable lowering of one transform for a fixed input size.

    ``src``: T describing the input (its buffer may be supplied by the caller so that several
    producers fill channel slices of one pixel-interleaved tensor).  ``dst_into``: optional
    (Buffer, c_off) the last stage writes into (concat for free)."""

    def __init__(self, module, h, w, cin, device, cfg=DEFAULT, src_buf=None, src_c_off=0,
                 dst_into=None, out_scale=None, out_post='none', in_dtype=None, in_embed=None,
                 pad_cout=0, out_dtype=None, out_pad=0, in_exact=False):
        """in_embed=(buffer_channels, offset, weight_scale): the module's `cin` input channels
        are channels [offset, offset+cin) of a wider (zero-padded) pixel of `buffer_channels`
        channels; the first stage's weights are embedded accordingly and scaled.
        pad_cout: round the last stage's output channels up to this multiple (extra channels
        get zero weights), so narrow pixel-domain outputs still fill a tensor-core tile.
        in_exact: every input channel holds values that are exact in bf16 (8-bit level units): in the split-bf16 mode
        the stages reading the input (through the space-to-depth repack) skip the lo.Whi third of their MMAs."""
        self.cfg, self.device = cfg, torch.device(device)
        self._graphs, self._warm, self._in_exact = {}, set(), True
        g = Graph()
        self.src = T(h, w, cin, external=True)
        self.dst = lower(module, self.src, g)
        self.stages = g.stages
        if not self.stages:
            raise ValueError('nothing to run')
        if in_embed is not None:
            buf_c, off, wsc = in_embed
            for s in self.stages:
                if s.src is self.src:
                    s.cin_off, s.w_scale = off, wsc
            self.src.c = buf_c
        if cfg.tc:
            self._split_narrow_tconvs()
            if in_embed is not None and cfg.s2d_first:
                self._space_to_depth_first_layer()
            if in_exact and cfg.precision == 'bf16x3':
                first = {id(s.dst) for s in self.stages if s.kind == 3 and s.src is self.src}
                for s in self.stages:
                    if s.kind == 0 and (s.src is self.src or id(s.src) in first):
                        s.flags |= _lib.OP_IN_EXACT
        last = self.stages[-1]
        self.out_c = self.dst.c
        if pad_cout and self.dst.c % pad_cout and last.kind != 2:
            assert last.gdn is None and last.res is None and last.gate is None
            self.dst.c = -(-self.dst.c // pad_cout) * pad_cout
        if out_scale is not None:
            last.out_scale = out_scale.detach().float().reshape(-1)
        if out_post != 'none':
            assert last.post == 'none'
            last.post = out_post
        self._choose_engines()
        if out_dtype is not None and dst_into is None:     # e.g. a bf16, bordered output for a tensor-core consumer
            self.dst.dtype, self.dst.pad = out_dtype, max(self.dst.pad, out_pad)
        self._assign_buffers(src_buf, src_c_off, dst_into, in_dtype)
        self._materialize()

    def _split_narrow_tconvs(self):
        """ConvTranspose2d with a handful of output channels (the 128 -> 3/6 pixel-domain ends):
        run the contraction as ONE 1x1 tensor-core stage producing k*k*cout channels per input
        pixel, then scatter them to the 2x output (col2im).  Reads the wide input once instead
        of once per tap."""
        out = []
        for s in self.stages:
            cin, cout = s.weight.shape[0], s.weight.shape[1]
            if (s.kind == 1 and cout <= 8 and s.k * s.k * cout <= 256 and cin % 16 == 0 and s.gdn is None
                    and s.res is None and s.gate is None and s.post == 'none' and s.out_scale is None
                    and s.cin_off == 0 and s.w_scale == 1.0):
                kk = s.k * s.k * cout
                wp = s.weight.permute(2, 3, 1, 0).reshape(kk, cin, 1, 1).contiguous()
                p_t = T(s.src.h, s.src.w, -(-kk // 32) * 32)       # (32-channel chunks of the persistent 1x1 kernel)
                out.append(Stage(0, 1, 1, s.src, p_t, wp, None, 'no'))
                out.append(Stage(2, s.k, 2, p_t, s.dst, None, s.bias, s.act))
            else:
                out.append(s)
        self.stages = out

    def _space_to_depth_first_layer(self):
        """5x5 stride-2 conv on the 16-channel pixel buffer (CustomConvLayer(5, in, C, stride 2), first
        layer of g_a / g_a_ref)  ==  3x3 stride-1 conv over the space-to-depth image (2x2 pixel blocks as
        64 channels): tap (ky, kx) = (2 ty + dy, 2 tx + dx) of the 5x5 kernel becomes channel block
        (dy, dx) of tap (ty, tx); the 11 positions of the 6x6 footprint outside the 5x5 get zero weights.
        K grows from 400 to 576, but the stage then runs on the persistent 3x3 kernel (one activation patch
        for all taps, fused GDN) instead of 25 strided 32-byte-per-pixel TMA gathers."""
        out = []
        for s in self.stages:
            w = s.weight
            if not (s.src is self.src and s.kind == 0 and s.k == 5 and s.stride == 2 and self.src.c == 16
                    and w.shape[0] in (64, 128) and s.gate is None):
                out.append(s)
                continue
            cout, cin_w = w.shape[0], w.shape[1]
            w16 = torch.zeros(cout, 16, 5, 5, dtype=torch.float32)
            w16[:, s.cin_off:s.cin_off + cin_w] = w.float().cpu() * s.w_scale
            w3 = s2d_weights(w16)
            t = T(s.dst.h, s.dst.w, 64)
            out.append(Stage(3, 1, 1, s.src, t, None, None))
            px = s.dst.h * s.dst.w
            flops = 2.0 * 25 * cin_w * cout * px + (2.0 * cout * cout * px if s.gdn is not None else 0.0)
            out.append(Stage(0, 3, 1, t, s.dst, w3, s.bias, s.act, s.post, s.gdn, s.res, None, s.out_scale,
                             alg_flops=flops))
        self.stages = out

    # -- engine / dtype / border selection
    def _choose_engines(self):
        tc, x3 = self.cfg.tc, self.cfg.precision == 'bf16x3'
        for s in self.stages:
            if s.kind == 3:
                s.engine = ENGINE_SIMT
                s.src.dtype = s.dst.dtype = self.cfg.act_dtype
                continue
            if s.kind == 2:
                s.engine = ENGINE_SIMT
                if tc and not x3:   # partial sums of the pixel-domain output: fp16 (11-bit mantissa), half the bytes of fp32
                    s.src.dtype = F16
                continue
            cin, cout = s.src.c, s.dst.c
            ok = tc and cin % self.cfg.tc_min_cin == 0 and cout % 16 == 0 and cout <= 256
            if s.gdn is not None and cout > 128:
                ok = False
            s.engine = (ENGINE_TC_X3 if x3 else ENGINE_TC) if ok else ENGINE_SIMT
        for s in self.stages:
            if s.engine != ENGINE_SIMT:
                assert s.src.dtype != F16
                s.src.dtype = self.cfg.act_dtype
                if s.kind == 0 and s.k > 1:
                    s.src.pad = max(s.src.pad, s.k // 2)
                # skip / trunk tensors that only ever feed an epilogue (ChengResBlock aux path, attention
                # trunk) are bf16 too: their producers then store through the staged TMA path and the
                # consumer prefetches them as packed 16-byte groups
                for t in (s.res, s.gate):
                    if t is not None and not t.external:
                        t.dtype = self.cfg.act_dtype

    def _assign_buffers(self, src_buf, src_c_off, dst_into, in_dtype):
        for i, s in enumerate(self.stages):
            for t in (s.src, s.res, s.gate):
                if t is not None:
                    t.last = max(t.last, i)
            s.dst.first = i
            s.dst.last = max(s.dst.last, i)
        if src_buf is not None:
            self.src.buf, self.src.c_off = src_buf, src_c_off
            if self.src.dtype != src_buf.dtype or src_buf.pad < self.src.pad:
                raise ValueError('supplied input buffer has dtype %d pad %d, plan needs dtype %d pad %d'
                                 % (src_buf.dtype, src_buf.pad, self.src.dtype, self.src.pad))
        else:
            if in_dtype is not None and self.src.dtype == F32:
                self.src.dtype = in_dtype
            self.src.buf = Buffer(self.src.h, self.src.w, self.src.c, self.src.pad, self.src.dtype,
                                  self.device)
        self.dst.last = 10 ** 9
        if dst_into is not None:
            buf, off = dst_into
            if (buf.h, buf.w) != (self.dst.h, self.dst.w):
                raise ValueError('dst_into size mismatch')
            self.dst.buf, self.dst.c_off, self.dst.dtype, self.dst.pad = buf, off, buf.dtype, buf.pad
        free, self.buffers = {}, []
        for i, s in enumerate(self.stages):
            t = s.dst
            lane = s.flags & _lib.OP_LANE1
            if t.buf is None:
                key = (t.h, t.w, t.c, t.pad, t.dtype, lane)
                pool = free.setdefault(key, [])
                t.buf = pool.pop() if pool else Buffer(t.h, t.w, t.c, t.pad, t.dtype, self.device)
                if t.buf not in self.buffers:
                    self.buffers.append(t.buf)
            # release tensors whose last reader is this stage (never the one just written)
            for u in {id(x): x for x in (s.src, s.res, s.gate) if x is not None}.values():
                if u.last == i and not u.external and u is not self.dst and u.buf is not None \
                        and u.c_off == 0 and u.c == u.buf.c:
                    # recycled only within the lane of the stage that read it last (streams of the
                    # two lanes are ordered against each other only at fork / join)
                    free.setdefault((u.h, u.w, u.c, u.pad, u.dtype, lane), []).append(u.buf)

    def _materialize(self):
        L = _lib.lib()
        dev = self.device
        n = len(self.stages)
        self.ops = (ConvOp * n)()
        self._keep = []
        st = _lib.stream_ptr()
        for i, s in enumerate(self.stages):
            op = self.ops[i]
            cin, cout = s.src.c, s.dst.c
            op.kind, op.k, op.stride, op.engine = s.kind, s.k, s.stride, s.engine
            op.flags = s.flags if self.cfg.two_lanes else (s.flags & _lib.OP_IN_EXACT)
            op.inp, op.out = s.src.fmap(), s.dst.fmap()
            op.alg_flops = s.alg_flops
            if s.kind == 3:
                continue
            if s.kind == 2:
                op.act = ACT[s.act]
                if s.bias is not None:
                    b = s.bias.to(dev, torch.float32).contiguous()
                    op.bias = b.data_ptr()
                    self._keep.append(b)
                continue
            wsrc = s.weight.to(dev, torch.float32).contiguous()
            w_cout, w_cin = (wsrc.shape[0], wsrc.shape[1]) if s.kind == 0 else (wsrc.shape[1], wsrc.shape[0])
            nbytes = L.aivc_packed_weight_bytes(s.k, s.engine, cin, cout)
            wdst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.check(L.aivc_pack_conv_weight(wsrc.data_ptr(), wdst.data_ptr(), s.kind, s.k, w_cin, w_cout,
                                               s.engine, cin, cout, s.cin_off, s.w_scale, st))
            op.weight = wdst.data_ptr()
            self._keep += [wsrc, wdst]
            if s.bias is not None:
                b = torch.zeros(cout, dtype=torch.float32, device=dev)
                b[:w_cout] = s.bias.to(dev, torch.float32)
                op.bias = b.data_ptr()
                self._keep.append(b)
            op.act, op.post = ACT[s.act], POST[s.post]
            if s.gdn is not None:
                beta, gamma, _ = s.gdn
                beta = beta.to(dev).contiguous()
                if s.engine == ENGINE_SIMT:
                    gm = gamma.to(dev).t().contiguous()                 # [j][i]
                    scratch = torch.empty(cout * s.dst.h * s.dst.w, dtype=torch.float32, device=dev)
                    op.scratch = scratch.data_ptr()
                    self._keep.append(scratch)
                elif s.engine == ENGINE_TC_X3:
                    g_hi = gamma.to(dev).to(torch.bfloat16)             # [i][hi j | lo j]
                    g_lo = (gamma.to(dev) - g_hi.float()).to(torch.bfloat16)
                    gm = torch.cat([g_hi, g_lo], dim=1).contiguous()
                else:
                    gm = gamma.to(dev).to(torch.bfloat16).contiguous()  # [i][j], K-major B operand
                op.gdn_beta, op.gdn_gamma = beta.data_ptr(), gm.data_ptr()
                self._keep += [beta, gm]
            if s.res is not None:
                op.residual = s.res.fmap()
            if s.gate is not None:
                op.gate = s.gate.fmap()
            if s.out_scale is not None:
                sc = torch.ones(cout, dtype=torch.float32, device=dev)
                sc[:s.out_scale.numel()] = s.out_scale.to(dev, torch.float32)
                op.out_scale = sc.data_ptr()
                self._keep.append(sc)
        torch.cuda.current_stream().synchronize()

    @property
    def in_fmap(self):
        return self.src.fmap()

    @property
    def out_fmap(self):
        fm = self.dst.fmap()
        fm.c = self.out_c              # hide the zero-weight padding channels
        return fm

    def run(self):
        """Enqueue the transform on the current stream.  With cfg.cuda_graphs the stages are replayed as ONE CUDA graph
        (captured on the second call, after a plain first run has set the kernels' one-time attributes); a graph is
        keyed by the pointers a caller may swap between runs (the last stage's out_scale = the per-frame-type gain)."""
        L = _lib.lib()
        if not self.cfg.cuda_graphs or _lib.PROFILING:
            _lib.check(L.aivc_conv2d_fused_seq(self.ops, len(self.ops), _lib.stream_ptr()))
            return
        key = (self.ops[len(self.ops) - 1].out_scale, self._in_exact)
        g = self._graphs.get(key)
        if g is None:
            if key not in self._warm:
                self._warm.add(key)
                _lib.check(L.aivc_conv2d_fused_seq(self.ops, len(self.ops), _lib.stream_ptr()))
                return
            ex = C.c_void_p()
            _lib.check(L.aivc_plan_graph_create(self.ops, len(self.ops), C.byref(ex)))
            g = self._graphs[key] = ex
        _lib.check(L.aivc_plan_graph_launch(g, _lib.stream_ptr()))

    def set_input_exact(self, exact):
        """Stages flagged OP_IN_EXACT (see `in_exact`) skip the lo halves of their input only while the caller vouches
        that the input is exact in bf16 (uint8 planes); float planes switch the full three-part product back on."""
        exact = bool(exact)
        if exact != self._in_exact:
            for op, s in zip(self.ops, self.stages):
                if s.flags & _lib.OP_IN_EXACT:
                    op.flags = (op.flags | _lib.OP_IN_EXACT) if exact else (op.flags & ~_lib.OP_IN_EXACT)
            self._in_exact = exact

    def __del__(self):
        try:
            L = _lib.lib()
            for g in getattr(self, '_graphs', {}).values():
                L.aivc_plan_graph_destroy(g)
        except Exception:
            pass

    def flops(self):
        """Algorithmic FLOPs of the reference graph: convs, tconvs and GDN 1x1."""
        f = 0
        for s in self.stages:
            if s.kind >= 2:
                continue
            if s.alg_flops:
                f += s.alg_flops
                continue
            px = s.dst.h * s.dst.w if s.kind == 0 else s.src.h * s.src.w
            w_cout, w_cin = (s.weight.shape[0], s.weight.shape[1]) if s.kind == 0 \
                else (s.weight.shape[1], s.weight.shape[0])
            f += 2 * s.k * s.k * w_cin * w_cout * px
            if s.gdn is not None:
                f += 2 * s.dst.c * s.dst.c * s.dst.h * s.dst.w
        return f


# ------------------------------------------------------------------ nn.Module boundary
# Plans of a module, per input geometry.  Kept OUT of the module (a WeakKeyDictionary keyed by the module object): the
# reference stores whole-module pickles, and `torch.save(model)` / `copy.deepcopy(model)` must not meet ctypes arrays
# with device pointers in `module.__dict__`; the entry dies with the module, so a recycled id() cannot alias it.
_PLANS = weakref.WeakKeyDictionary()


def cached_plans(module):
    return _PLANS.setdefault(module, {})


def weights_version(module):
    """Changes whenever a parameter or buffer of `module` is modified in place (load_state_dict, optimiser step,
    manual edits under no_grad): packed weights of a cached plan are then stale."""
    v = 0
    for t in module.parameters():
        v += t._version
    for t in module.buffers():
        v += t._version
    return v


def set_default_config(cfg):
    """Config used by the drop-in modules' forward() (module(x)) from now on; returns the previous one.  The default is
    Config() = the fp32-faithful split-bf16 tensor-core mode."""
    global DEFAULT
    old, DEFAULT = DEFAULT, cfg
    return old


def run_module(module, x, cfg=None):
    """forward() of the drop-in classes: NCHW fp32 CUDA tensor in, NCHW fp32 tensor out."""
    cfg = DEFAULT if cfg is None else cfg
    if not (isinstance(x, torch.Tensor) and x.is_cuda):
        raise RuntimeError('aivc_b200 layers run on a CUDA device only (no CPU fallback); got '
                           + str(getattr(x, 'device', type(x))))
    if x.dim() != 4 or x.shape[0] != 1:
        raise ValueError('expected a [1, C, H, W] tensor (AIVC codes one frame at a time)')
    L = _lib.lib()
    x = x.contiguous().float()
    _, c, h, w = x.shape
    cache = cached_plans(module)
    key = (h, w, c, x.device.index, cfg.key())
    ver = weights_version(module)
    hit = cache.get(key)
    if hit is None or hit[0] != ver:              # (weights changed since the plan packed them: lower again)
        with torch.cuda.device(x.device):
            hit = cache[key] = (ver, Plan(module, h, w, c, x.device, cfg))
    plan = hit[1]
    with torch.cuda.device(x.device):
        st = _lib.stream_ptr()
        fin = plan.in_fmap
        _lib.check(L.aivc_nchw_to_fmap(x.data_ptr(), C.byref(fin), st))
        plan.run()
        fout = plan.out_fmap
        out = torch.empty((1, fout.c, fout.h, fout.w), dtype=torch.float32, device=x.device)
        _lib.check(L.aivc_fmap_to_nchw(C.byref(fout), out.data_ptr(), st))
    return out
