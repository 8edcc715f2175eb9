"""GPU suite for the tcgen05 (bf16 operands, fp32 accumulate) engine.

Tolerance: operands are rounded to bf16 (relative 2^-9) before every contraction and
activations are stored as bf16 between stages, so against the fp32 oracle we require
  rms(y - ref) <= 1% of rms(ref) per fused stage (3% for multi-stage blocks)
  max|y - ref| <= 6% of max|ref|
which a wrong tap, channel, swizzle or border (errors of order 100%) cannot meet.
"""
import numpy as np
import pytest
import torch

from tests.leafcfg import LEAVES, load_leaf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    return torch.device('cuda:0')


def _cfg():
    from aivc_b200.plan import Config
    return Config(precision='bf16')


def _check(y, ref, rms_tol, max_tol=0.06):
    err = y - ref
    rms = np.sqrt((err ** 2).mean()) / max(np.sqrt((ref ** 2).mean()), 1e-12)
    mx = np.abs(err).max() / max(np.abs(ref).max(), 1e-12)
    assert rms <= rms_tol and mx <= max_tol, 'rms %.4f (tol %.3f), max %.4f (tol %.3f)' % (rms, rms_tol, mx, max_tol)


def _uses_tc(m, x):
    from aivc_b200 import plan
    from aivc_b200._lib import ENGINE_TC
    p = next(iter(plan.cached_plans(m).values()))[1]
    return sum(1 for s in p.stages if s.engine == ENGINE_TC), len(p.stages)


@pytest.mark.parametrize('name', sorted(LEAVES))
def test_leaf_bf16_vs_reference_golden(name, golden_dir, dev):
    from aivc_b200 import plan
    m, fx = load_leaf(name, golden_dir)
    for i in range(2):
        y = plan.run_module(m, torch.from_numpy(fx['x%d' % i]).to(dev), _cfg()).cpu().numpy()
        _check(y, fx['y%d' % i], 0.03)


WIDE = {
    'conv3_s1_leaky_128': (lambda L: L.CustomConvLayer(3, 128, 128, non_linearity='leaky_relu'), 128, 0.01),
    'conv3_s2_no_128_64': (lambda L: L.CustomConvLayer(3, 128, 64, non_linearity='no', conv_stride=2), 128, 0.01),
    'conv3_s1_gdn_128': (lambda L: L.CustomConvLayer(3, 128, 128, non_linearity='gdn'), 128, 0.01),
    'conv3_s1_igdn_64': (lambda L: L.CustomConvLayer(3, 64, 64, non_linearity='gdn_inverse'), 64, 0.01),
    'conv5_s2_gdn_64_128': (lambda L: L.CustomConvLayer(5, 64, 128, non_linearity='gdn', conv_stride=2), 64, 0.01),
    'conv3_s1_relu_32_48': (lambda L: L.CustomConvLayer(3, 32, 48, non_linearity='relu'), 32, 0.01),
    'up3_no_128': (lambda L: L.UpscalingLayer(3, 128, 128, non_linearity='no'), 128, 0.01),
    'up5_leaky_64_32': (lambda L: L.UpscalingLayer(5, 64, 32, non_linearity='leaky_relu'), 64, 0.01),
    'cheng_plain_128': (lambda L: L.ChengResBlock(128, 'plain'), 128, 0.03),
    'cheng_down_128': (lambda L: L.ChengResBlock(128, 'down'), 128, 0.03),
    'cheng_up_64': (lambda L: L.ChengResBlock(64, 'up_tconv'), 64, 0.03),
    'attention_64': (lambda L: L.SimplifiedAttention(64), 64, 0.03),
    'attention_light_128': (lambda L: L.SimplifiedAttention(128, lightweight_resblock=True), 128, 0.03),
}


@pytest.mark.parametrize('name', sorted(WIDE))
@pytest.mark.parametrize('size', [(33, 47), (16, 128)])
def test_wide_layers_vs_oracle(name, size, dev):
    """Channel counts that exercise the 128-byte-swizzle K chunks, partial tiles (odd sizes)
    and the chained GDN GEMM; oracle = CPU fp32 restatement on the same seeded weights."""
    import aivc_b200.layers as M
    from aivc_b200 import plan
    from oracle import nn_ref as R
    mk, cin, tol = WIDE[name]
    torch.manual_seed(hash(name) % 1000)
    m = mk(M).eval()
    with torch.no_grad():
        for n_, p in m.named_parameters():
            if n_.endswith('gamma'):
                p.add_(0.02 * torch.rand_like(p))
    h, w = size
    x = torch.randn(1, cin, h, w, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        ref = R.forward_module(m, x).numpy()
    y = plan.run_module(m, x.to(dev), _cfg()).cpu().numpy()
    n_tc, n_all = _uses_tc(m, x)
    assert n_tc == n_all, 'expected every stage on the tcgen05 engine (%d of %d)' % (n_tc, n_all)
    assert y.shape == ref.shape
    _check(y, ref, tol)


@pytest.mark.parametrize('name', ['conv3_s1_leaky_128', 'cheng_plain_128', 'attention_64', 'conv3_s1_relu_32_48',
                                  'conv3_s1_gdn_128', 'cheng_down_128', 'up3_no_128', 'cheng_up_64'])
def test_persistent_3x3_kernel_vs_oracle(name, dev):
    """Maps large enough (>= 120 tiles of 32x8) for the persistent 3x3 kernel: odd sizes, partial
    tiles on both edges, residual / gate / post-activation epilogues, several tiles per CTA."""
    import aivc_b200.layers as M
    from aivc_b200 import plan
    from oracle import nn_ref as R
    mk, cin, tol = WIDE[name]
    torch.manual_seed(11)
    m = mk(M).eval()
    for h, w in ((135, 243), (270, 481)):
        x = torch.randn(1, cin, h, w, generator=torch.Generator().manual_seed(h))
        with torch.no_grad():
            ref = R.forward_module(m, x).numpy()
        y = plan.run_module(m, x.to(dev), _cfg()).cpu().numpy()
        _check(y, ref, tol)


def test_codec_bf16_closed_loop(golden_dir, dev):
    """bf16 engine on the golden system case: decoder reproduces the encoder bit for bit, and
    the reconstruction stays within a few 8-bit levels of the fp32 oracle."""
    import os
    from aivc_b200 import models, gop as G
    from aivc_b200.codec import FrameCodec, planes_to_device
    fx = np.load(os.path.join(golden_dir, 'system_80x112.npz'))
    h, w = int(fx['H']), int(fx['W'])
    net = models.build_standin(seed=4321, C=32, Cy=16, Cz=16, Csc=16)
    gop = G.generate_gop_struct('1_GOP_2')
    frames = {'frame_%d' % t: planes_to_device([fx['src_frame_%d_%s' % (t, k)] for k in 'yuv'], dev)
              for t in range(3)}
    codec = FrameCodec(net, h, w, dev, _cfg())
    bts, rec = codec.encode_gop(frames, gop)
    dec = codec.decode_gop(bts, gop)
    worst = 0
    for f in gop:
        for a, b in zip(rec[f], dec[f]):
            assert torch.equal(a, b)
        for k, p in zip('yuv', rec[f]):
            ref = fx['spec_rec_%s_%s' % (f, k)].reshape(-1).astype(np.int32)
            worst = max(worst, int(np.abs(p.cpu().numpy().astype(np.int32) - ref).max()))
    assert worst <= 8, 'bf16 reconstruction deviates by %d levels' % worst


@pytest.mark.parametrize('size', [(135, 241), (64, 96), (33, 47)])
@pytest.mark.parametrize('cin_ref, off', [(9, 0), (6, 3)])
def test_first_layer_space_to_depth_vs_oracle(size, cin_ref, off, dev):
    """The bf16 engine's first layer -- 5x5 stride-2 conv + GDN on the 16-channel level-unit pixel buffer,
    lowered to space-to-depth + 3x3 on the fused conv+GDN kernel -- against the oracle on odd and even sizes
    (ceil-halving, per-pixel replicate border), for g_a (channels 0..8) and g_a_ref (a slice at offset 3)."""
    import ctypes as C
    import aivc_b200.layers as M
    from aivc_b200 import _lib
    from aivc_b200.plan import Plan, Buffer, Config
    from aivc_b200._lib import BF16
    from oracle import nn_ref as R
    h, w = size
    torch.manual_seed(h + cin_ref)
    m = M.CustomConvLayer(5, cin_ref, 128, non_linearity='gdn', conv_stride=2).eval()
    levels = torch.randint(0, 256, (1, 9, h, w), generator=torch.Generator().manual_seed(w)).float()
    with torch.no_grad():
        ref = R.forward_module(m, levels[:, off:off + cin_ref] / 255.).numpy()
    L = _lib.lib()
    with torch.cuda.device(dev):
        buf = Buffer(h, w, 16, 2, BF16, dev)
        fm = buf.view(0, 9)
        xd = levels.to(dev).contiguous()
        _lib.check(L.aivc_nchw_to_fmap(xd.data_ptr(), C.byref(fm), _lib.stream_ptr()))
        plan = Plan(m, h, w, cin_ref, dev, Config(precision='bf16'), src_buf=buf, in_embed=(16, off, 1.0 / 255.0))
        assert [s.kind for s in plan.stages] == [3, 0] and plan.stages[1].k == 3 and plan.stages[1].src.c == 64
        plan.run()
        fo = plan.out_fmap
        out = torch.empty((1, fo.c, fo.h, fo.w), dtype=torch.float32, device=dev)
        _lib.check(L.aivc_fmap_to_nchw(C.byref(fo), out.data_ptr(), _lib.stream_ptr()))
    y = out.cpu().numpy()
    assert y.shape == ref.shape
    _check(y, ref, 0.01)
