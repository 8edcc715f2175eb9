"""CPU suite: host-side product logic (range coder, framing, GOP schedules, lowering) and the
C-ABI surface.  No compute kernel is launched."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from aivc_b200 import _lib, entropy, gop as G, plan as P
from oracle import codec_ref as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'aivc_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(aivc_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    L = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), 'missing export ' + name
    assert declared == set(_lib.EXPORTS)
    assert _lib.lib().aivc_abi_version() == 1


def _enc_bounds(bounds):
    L = _lib.lib()
    out = np.empty(L.aivc_rc_bound(bounds.size), np.uint8)
    n = C.c_size_t()
    _lib.check(L.aivc_rc_encode_bounds(bounds.ctypes.data, bounds.size, out.ctypes.data, out.size, C.byref(n)))
    return out[:n.value].tobytes()


def test_rangecoder_known_answers():
    """Hand-derived: with a uniform 2-symbol CDF {0, 32768, 65536} every symbol is one bit
    (0 -> '0', 1 -> '1'), followed by the terminating '01' / '10' pattern and zero padding."""
    half = np.uint32(0 | (32768 << 16))          # symbol 0: [0, 32768)
    # eight zeros: after 8 settled '0' bits low = 0 -> final bit 0 + one pending 1  => 00000000 01
    assert _enc_bounds(np.array([half] * 8, np.uint32)) == bytes([0x00, 0x40])
    # empty message: low = 0 -> '0' then pending '1'
    assert _enc_bounds(np.zeros(0, np.uint32)) == bytes([0x40])
    # oracle coder agrees on both
    assert O.rc_encode_bounds([0] * 8, [32768] * 8) == bytes([0x00, 0x40])
    assert O.rc_encode_bounds([], []) == bytes([0x40])


@pytest.mark.parametrize('seed', range(6))
def test_rangecoder_matches_oracle_and_roundtrips(seed):
    rng = np.random.default_rng(seed)
    L = _lib.lib()
    n = int(rng.integers(1, 4000))
    sig = np.exp(rng.uniform(-7, 4.5, n)).astype(np.float32)
    q = np.clip(np.rint(rng.laplace(0, sig / np.sqrt(2))), -256, 255).astype(np.int16)
    if seed == 0:
        q[:] = 0
    if seed == 1:
        q[::7] = 255
        q[3::7] = -256
    table = O.laplace_table_spec(sig)
    ar = np.arange(n)
    lo = table[ar, q.astype(int) + 256].astype(np.uint32)
    hi = table[ar, q.astype(int) + 257].astype(np.uint32)
    ours = _enc_bounds(lo | (hi << 16))
    assert ours == O.rc_encode_table(table, (q + 256).astype(np.int16))
    b = (sig / np.float32(1.41421354)).astype(np.float32)
    dec = np.empty(n, np.int16)
    buf = np.frombuffer(ours, np.uint8)
    _lib.check(L.aivc_rc_decode_laplace(b.ctypes.data, buf.ctypes.data, len(ours), n, dec.ctypes.data))
    assert np.array_equal(dec, q)
    assert np.array_equal(O.rc_decode_table(table, ours, n) - 256, q)
    # the windowed decoder (CDF entries 253..260 handed over by the device) gives the same symbols
    win = np.ascontiguousarray(table[:, 253:261]).astype(np.uint16)
    dec2 = np.empty(n, np.int16)
    _lib.check(L.aivc_rc_decode_laplace_win(b.ctypes.data, win.ctypes.data, buf.ctypes.data, len(ours), n,
                                            dec2.ctypes.data))
    assert np.array_equal(dec2, q)


def test_host_cdf_equals_oracle_spec():
    rng = np.random.default_rng(3)
    L = _lib.lib()
    sig = np.exp(rng.uniform(-9.3, 5.1, 300)).astype(np.float32)
    table = O.laplace_table_spec(sig)
    b = (sig / np.float32(1.41421354)).astype(np.float32)
    for j in range(sig.size):
        for i in (0, 1, 128, 250, 255, 256, 257, 258, 300, 512, 513):
            assert L.aivc_laplace_cdf_int_host(float(b[j]), i) == table[j, i]


def test_sigma_host_close_to_torch():
    v = torch.linspace(-25, 15, 4001)
    ref = torch.exp(0.5 * torch.clamp(v, -18.4207, 10.0)).numpy()
    L = _lib.lib()
    got = np.array([L.aivc_sigma_from_logvar_host(float(x)) for x in v.numpy()], np.float32)
    assert np.max(np.abs(got - ref) / ref) < 2.5e-7        # <= 2 ulp of torch's vectorised expf


def test_latent_framing_roundtrip():
    rng = np.random.default_rng(5)
    c, h, w = 8, 5, 7
    sig = np.exp(rng.uniform(-1, 2, (c, h, w))).astype(np.float32)
    q = np.clip(np.rint(rng.laplace(0, sig)), -256, 255).astype(np.int16)
    q[[1, 4]] = 0
    b = (sig / np.float32(1.41421354)).astype(np.float32)
    L = _lib.lib()
    bounds = np.empty((c, h * w), np.uint32)
    for ch in range(c):
        for i in range(h * w):
            s = int(q[ch].reshape(-1)[i]) + 256
            bb = float(b[ch].reshape(-1)[i])
            bounds[ch, i] = L.aivc_laplace_cdf_int_host(bb, s) | (L.aivc_laplace_cdf_int_host(bb, s + 1) << 16)
    nz = (np.abs(q).reshape(c, -1).sum(1) != 0).astype(np.int32)
    sec = entropy.encode_y(bounds, nz)
    n = int.from_bytes(sec[:4], 'big')
    assert n == len(sec) - 4 and sec[4] == 6 and list(sec[5:11]) == [0, 2, 3, 5, 6, 7]
    assert np.array_equal(entropy.decode_y(sec[4:], b, c, h, w), q)
    # same bytes as the oracle's framing of the same latent
    ref = O.ac_encode_latent(torch.from_numpy(q.astype(np.float32))[None], 'laplace',
                             sigma=torch.from_numpy(sig)[None], cdf_mode='spec')
    assert ref == sec
    # z / pmf mode
    from aivc_b200.layers import BallePdfEstim
    torch.manual_seed(0)
    pz = BallePdfEstim(4, '')
    tab = entropy.z_table_u16(pz)
    assert np.array_equal(tab, O.z_table_u16(pz))
    z = rng.integers(-6, 7, (4, 3, 2)).astype(np.int16)
    sz = entropy.encode_z(tab, z)
    assert sz == O.ac_encode_latent(torch.from_numpy(z.astype(np.float32))[None], 'pmf', z_table=tab)
    assert np.array_equal(entropy.decode_z(tab, sz[4:], 4, 3, 2), z)
    secs = entropy.split_sections((0).to_bytes(4, 'big') * 2 + sz + sec)
    assert secs[0] == b'' and secs[1] == b'' and secs[2] == sz[4:] and secs[3] == sec[4:]


def test_gop_levels():
    g = G.generate_gop_struct('1_GOP_32')
    assert [len(l) for l in G.levels(g)] == [1, 1, 1, 2, 4, 8, 16]
    assert G.coding_order(g)[:7] == ['frame_%d' % i for i in (0, 32, 16, 8, 4, 2, 1)]
    g = G.generate_gop_struct('LDP_8')
    assert [len(l) for l in G.levels(g)] == [1] * 9


def test_lowering_fuses_blocks():
    import aivc_b200.layers as M
    g = P.Graph()
    x = P.T(32, 48, 16, external=True)
    y = P.lower(M.ChengResBlock(16, 'down'), x, g)
    assert (y.h, y.w, y.c) == (16, 24, 16)
    assert [(s.k, s.stride, s.act) for s in g.stages] == [(1, 2, 'no'), (3, 2, 'leaky_relu'), (3, 1, 'gdn')]
    assert g.stages[2].res is g.stages[0].dst
    g = P.Graph()
    y = P.lower(M.SimplifiedAttention(16), x, g)
    assert len(g.stages) == 13
    last = g.stages[-1]
    assert last.act == 'sigmoid' and last.gate is g.stages[5].dst and last.res is x
    assert all(s.post == 'relu' for i, s in enumerate(g.stages[:12]) if i % 2 == 1)
    g = P.Graph()
    y = P.lower(M.ChengResBlock(16, 'up_tconv'), x, g)
    assert (y.h, y.w) == (64, 96) and [s.kind for s in g.stages] == [1, 1, 0]


def test_layers_refuse_cpu():
    import aivc_b200.layers as M
    with pytest.raises(RuntimeError):
        M.CustomConvLayer(3, 4, 4)(torch.zeros(1, 4, 8, 8))


@pytest.mark.parametrize('size', [(12, 16), (13, 17), (9, 10)])
def test_space_to_depth_first_layer_is_the_same_convolution(size):
    """plan.s2d_weights: replicate-pad(2) + 5x5 stride-2 conv == 3x3 stride-1 conv over the space-to-depth
    image with per-pixel replicate semantics (what the kind-3 stage + persistent kernel compute)."""
    import torch
    import torch.nn.functional as F
    from aivc_b200.plan import s2d_weights
    h, w = size
    g = torch.Generator().manual_seed(h * 100 + w)
    x = torch.randn(1, 16, h, w, generator=g, dtype=torch.float64)
    w16 = torch.randn(8, 16, 5, 5, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.pad(x, (2, 2, 2, 2), mode='replicate'), w16, stride=2)
    hb, wb = (h + 1) // 2, (w + 1) // 2
    ys = torch.arange(-1, hb + 1)[:, None] * 2 + torch.arange(2)[None, :]          # [hb+2, dy]
    xs = torch.arange(-1, wb + 1)[:, None] * 2 + torch.arange(2)[None, :]          # [wb+2, dx]
    ys, xs = ys.clamp(0, h - 1), xs.clamp(0, w - 1)
    blocks = x[0][:, ys][:, :, :, xs]                                              # [c, hb+2, dy, wb+2, dx]
    s2d = blocks.permute(2, 4, 0, 1, 3).reshape(1, 64, hb + 2, wb + 2)             # channel = (dy*2+dx)*16 + c
    out = F.conv2d(s2d, s2d_weights(w16.float()).double())
    assert out.shape == ref.shape
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-5)      # (weights pass through fp32)


def test_yuv_file_io_roundtrip_and_gop_schedule(tmp_path):
    """yuvio: name parser (format_conversion/utils.py:45-50, 69-72), frame layout, pinned staging, the
    reference's GOP split with a padded last GOP (model_management.py:142-173)."""
    import torch
    from aivc_b200 import yuvio
    assert yuvio.parse_name('/x/y/BlowingBubbles_416x240_50_420.yuv') == (416, 240, 50.0)
    with pytest.raises(ValueError):
        yuvio.parse_name('clip.yuv')
    w, h, n = 18, 10, 5
    rng = np.random.default_rng(3)
    frames = [(rng.integers(0, 256, w * h, dtype=np.uint8), rng.integers(0, 256, (w // 2) * (h // 2), dtype=np.uint8),
               rng.integers(0, 256, (w // 2) * (h // 2), dtype=np.uint8)) for _ in range(n)]
    path = str(tmp_path / ('clip_%dx%d_30_420.yuv' % (w, h)))
    with yuvio.YuvWriter(path) as wr:
        for f in frames:
            wr.append(f)
    rd = yuvio.YuvReader(path)
    assert (len(rd), rd.w, rd.h, rd.fps) == (n, w, h, 30.0) and rd.fb == w * h * 3 // 2
    for i in (0, 4, 2):
        for a, b, c in zip(rd.frame(i), frames[i], rd.pinned(i)):
            assert np.array_equal(a, b) and np.array_equal(c.numpy(), b)
    with pytest.raises(IndexError):
        rd.frame(n)
    assert yuvio.gop_schedule(0, 4, 3) == [(0, 3), (3, 2)] and yuvio.gop_schedule(2, 4, 3) == [(2, 3)]
    g = rd.gop_frames(3, 3, 4, torch.device('cpu'))           # frames 3, 4 and the padding copy of 4
    assert np.array_equal(g['frame_1'][0].numpy(), frames[4][0]) and np.array_equal(g['frame_2'][0].numpy(), frames[4][0])
    assert np.array_equal(g['frame_0'][2].numpy(), frames[3][2])
    open(str(tmp_path / 'bad_18x10_30_420.yuv'), 'wb').write(b'123')
    with pytest.raises(ValueError):
        yuvio.YuvReader(str(tmp_path / 'bad_18x10_30_420.yuv'))


def test_codec_cache_is_versioned_weak_and_bounded(monkeypatch):
    """adapter.codec_for: one codec per (size, device, config, rate) and model, rebuilt when the weights change,
    at most MAX_CODECS_PER_MODEL alive, gone with the model, nothing stored on the model itself."""
    import gc
    import torch
    from aivc_b200 import adapter, models
    built = []

    class Fake:
        def __init__(self, model, h, w, device, cfg, idx_rate):
            built.append((h, w))
    monkeypatch.setattr(adapter, 'FrameCodec', Fake)
    net = models.build_standin(seed=1, C=16, Cy=8, Cz=8, Csc=8)
    a = adapter.codec_for(net, 64, 64, 'cuda:0')
    assert adapter.codec_for(net, 64, 64, 'cuda:0') is a and len(built) == 1
    with torch.no_grad():
        next(net.parameters()).add_(1.0)
    b = adapter.codec_for(net, 64, 64, 'cuda:0')
    assert b is not a and len(built) == 2                    # weights changed -> packed weights are stale
    adapter.codec_for(net, 32, 32, 'cuda:0')
    adapter.codec_for(net, 16, 16, 'cuda:0')                 # third geometry: the least recently used one goes
    assert len(adapter._CODECS[net]) == adapter.MAX_CODECS_PER_MODEL
    assert adapter.codec_for(net, 64, 64, 'cuda:0') is not b
    assert not any(k.startswith('_aivc') for k in net.__dict__)
    n = len(adapter._CODECS)
    del net, a, b
    gc.collect()
    assert len(adapter._CODECS) == n - 1


@pytest.mark.parametrize('size', [(270, 480), (272, 480), (270, 481), (540, 960), (544, 960), (360, 640), (1080, 1920),
                                  (300, 1000), (33, 47), (600, 808)])
@pytest.mark.parametrize('sms', [148, 132, 7])
def test_persistent_3x3_tiling_covers_the_image_exactly_once(size, sms):
    """Host-side check of the work-item cut of the persistent split-bf16 3x3 kernel (csrc/conv_tc3.cu::choose_tiling +
    item_tile, the same function the device code runs): whole 32x8 tiles and 16-row half tiles together cover every
    pixel of the layer exactly once, the whole tiles fill complete waves when half tiles are used, and a mixed cut
    is only chosen where its estimated cost is below that of whole tiles."""
    h, w = size
    L = _lib.lib()
    cap = 1 << 16
    buf = np.zeros((cap, 3), dtype=np.int32)
    n = L.aivc_debug_tc3_tiling(h, w, sms, buf.ctypes.data_as(C.c_void_p), cap)
    assert 0 < n <= cap
    t = buf[:n]
    cover = np.zeros((h + 64, w + 16), dtype=np.int32)
    for y0, x0, rows in t:
        assert rows in (16, 32) and x0 % 8 == 0 and y0 % 16 == 0
        cover[y0:y0 + rows, x0:x0 + 8] += 1
    assert (cover[:h, :w] == 1).all(), 'every pixel exactly once'
    assert cover.max() == 1, 'no tile overlaps another, inside or outside the image'
    nfull, nhalf = int((t[:, 2] == 32).sum()), int((t[:, 2] == 16).sum())
    assert (t[:nfull, 2] == 32).all(), 'whole tiles first'
    whole_only = -(-w // 8) * -(-h // 32)
    if nhalf:
        assert nfull % sms == 0
        # cost model of choose_tiling: a half tile counts 0.75 of a whole one
        assert 2.0 * (nfull // sms) + 1.5 * -(-nhalf // sms) < 2.0 * -(-whole_only // sms)
    else:
        assert n == whole_only
    if size in ((270, 480), (272, 480)) and sms == 148:
        assert (nfull, nhalf) == (444, 132)
