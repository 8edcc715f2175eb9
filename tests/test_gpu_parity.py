"""GPU parity suite (B200): the CUDA path, called through the C ABI, against the golden
vectors the oracle pinned to the reference classes, and against the oracle itself.

Tolerances (fp32 engine): the kernels accumulate in fp32 like the reference's CPU path but in
a different order, so values agree to a few ulp of the accumulated magnitude: rtol 2e-5 /
atol 2e-5 on activations of order 1.  Integer results (symbols, CDF bounds, 8-bit planes,
bitstream bytes) must be exact on identical inputs.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from tests.leafcfg import LEAVES, load_leaf

pytestmark = pytest.mark.gpu

RTOL, ATOL = 2e-5, 2e-5


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU suite needs a CUDA device'
    return torch.device('cuda:0')


def _fp32():
    from aivc_b200.plan import Config
    return Config(precision='fp32')


@pytest.mark.parametrize('name', sorted(LEAVES))
def test_leaf_fp32_matches_reference_golden(name, golden_dir, dev):
    from aivc_b200 import plan
    m, fx = load_leaf(name, golden_dir)
    for i in range(2):
        x = torch.from_numpy(fx['x%d' % i]).to(dev)
        y = plan.run_module(m, x, _fp32()).cpu().numpy()
        ref = fx['y%d' % i]
        assert y.shape == ref.shape
        np.testing.assert_allclose(y, ref, rtol=RTOL, atol=ATOL)


def test_default_forward_is_the_cuda_path(golden_dir, dev):
    """module(x) itself (the drop-in call) runs the library, in the default precision."""
    m, fx = load_leaf('cheng_plain', golden_dir)
    y = m(torch.from_numpy(fx['x0']).to(dev)).cpu().numpy()
    np.testing.assert_allclose(y, fx['y0'], rtol=0.05, atol=0.05)


def test_pixel_ends(golden_dir, dev):
    from aivc_b200 import ops, _lib
    from aivc_b200.plan import Buffer
    from aivc_b200._lib import F32
    fx = np.load(os.path.join(golden_dir, 'misc.npz'))
    L = _lib.lib()
    for tag in ('even', 'odd'):
        yuv = [torch.from_numpy(fx['%s_in_%s' % (tag, k)]).to(dev) for k in 'yuv']
        x444 = ops.yuv420_to_444(*yuv)
        assert np.array_equal(x444.cpu().numpy(), fx[tag + '_x444'])       # exact: pure data movement
        # warp (stand-alone motion compensation with beta = 1)
        flow = torch.from_numpy(fx[tag + '_flow']).to(dev)
        beta = torch.ones_like(x444)
        wr = ops.warp_blend(x444, x444, flow, flow, beta).cpu().numpy()
        np.testing.assert_allclose(wr, fx[tag + '_warp'], rtol=1e-5, atol=2e-6)
        # finalize: (x*1.3 - 0.1) through OutputLayer + pad/crop + 8-bit cast, exact levels
        _, _, h, w = x444.shape
        src = (x444 * 1.3 - 0.1)[0].permute(1, 2, 0).contiguous()
        buf = Buffer(h, w, 3, 0, F32, dev)
        buf.interior().copy_(src)
        fm = buf.view()
        hc, wc = (h + 1) // 2, (w + 1) // 2
        yo = torch.empty(h * w, dtype=torch.uint8, device=dev)
        uo = torch.empty(hc * wc, dtype=torch.uint8, device=dev)
        vo = torch.empty(hc * wc, dtype=torch.uint8, device=dev)
        _lib.check(L.aivc_finalize_frame(C.byref(fm), None, yo.data_ptr(), uo.data_ptr(), vo.data_ptr(),
                                         None, _lib.stream_ptr()))
        for k, t, shp in (('y', yo, (h, w)), ('u', uo, (hc, wc)), ('v', vo, (hc, wc))):
            ref = np.rint(fx['%s_fin_%s' % (tag, k)] * 255).astype(np.uint8).reshape(shp)
            got = t.cpu().numpy().reshape(shp)
            # the source was computed on the GPU (x*1.3-0.1): allow a level flip at exact ties only
            assert np.abs(got.astype(int) - ref.astype(int)).max() <= 1
            assert (got != ref).mean() < 1e-3


def test_mu_sigma_and_cdf_bounds_exact(golden_dir, dev):
    """sigma and the 16-bit CDF bounds are integer/bit-exact against the oracle's C restatement."""
    from aivc_b200 import ops, _lib
    from aivc_b200.plan import Buffer
    from aivc_b200._lib import F32
    from oracle import codec_ref as O
    fx = np.load(os.path.join(golden_dir, 'misc.npz'))
    L = _lib.lib()
    hs = torch.from_numpy(fx['hs_out']).to(dev)
    mu, sigma = ops.mu_sigma(hs, 8)
    assert np.array_equal(mu.cpu().numpy(), fx['mu'])
    sig = sigma.cpu().numpy()
    np.testing.assert_allclose(sig, fx['sigma'], rtol=2.5e-7, atol=0)          # vs torch expf: <= 2 ulp
    host = np.array([L.aivc_sigma_from_logvar_host(float(v)) for v in fx['hs_out'][0, 8:].reshape(-1)],
                    np.float32)
    assert np.array_equal(sig.reshape(-1), host)                               # device == host, bitwise
    # quantise a latent against (mu, sigma): symbols, bounds, non-zero flags, dequantised y
    rng = np.random.default_rng(0)
    c, h, w = 8, 9, 11
    y = (fx['mu'][0] + rng.laplace(0, 2.0, (c, h, w))).astype(np.float32)
    y[[2, 5]] = fx['mu'][0][[2, 5]] + 0.2                                      # two all-zero channels
    ybuf = Buffer(h, w, c, 0, F32, dev)
    ybuf.interior().copy_(torch.from_numpy(y).permute(1, 2, 0))
    hsbuf = Buffer(h, w, 2 * c, 0, F32, dev)
    hsbuf.interior().copy_(hs[0].permute(1, 2, 0))
    yhat = Buffer(h, w, c, 1, F32, dev)
    gain = torch.linspace(0.5, 1.5, c, device=dev)
    q = torch.empty(c * h * w, dtype=torch.int16, device=dev)
    bounds = torch.empty(c * h * w, dtype=torch.int32, device=dev)
    nz = torch.zeros(c, dtype=torch.int32, device=dev)
    fy, fh, fo = ybuf.view(), hsbuf.view(), yhat.view()
    rate = torch.empty(c * h * w, dtype=torch.float32, device=dev)
    _lib.check(L.aivc_quantize_latent(C.byref(fy), C.byref(fh), gain.data_ptr(), q.data_ptr(),
                                      bounds.data_ptr(), nz.data_ptr(), C.byref(fo), rate.data_ptr(),
                                      _lib.stream_ptr()))
    qn = q.cpu().numpy().reshape(c, h, w)
    # rate estimate of every symbol (pdf_estimator.py:27-70 + entropy_coder.py:25-30) against torch's evaluation
    from oracle import nn_ref as R
    ref_rate = R.laplace_rate_bits(torch.from_numpy(qn.astype(np.float32))[None], torch.from_numpy(fx['sigma'])).numpy()
    np.testing.assert_allclose(rate.cpu().numpy().reshape(1, c, h, w), ref_rate, rtol=2e-5, atol=2e-5)
    ref_q = np.clip(np.rint(y - fx['mu'][0]), -256, 255).astype(np.int16)
    assert np.array_equal(qn, ref_q)
    assert list(nz.cpu().numpy()) == [1, 1, 0, 1, 1, 0, 1, 1]
    lo = np.empty(c * h * w, np.uint32)
    hi = np.empty(c * h * w, np.uint32)
    sflat = np.ascontiguousarray(sig.reshape(-1))
    qflat = np.ascontiguousarray(ref_q.reshape(-1))
    O.lib().laplace_spec_bounds(sflat.ctypes.data, qflat.ctypes.data, qflat.size, lo.ctypes.data, hi.ctypes.data)
    got = bounds.cpu().numpy().view(np.uint32)
    assert np.array_equal(got & 0xFFFF, lo) and np.array_equal(got >> 16, hi)
    ref_yhat = (ref_q.astype(np.float32) + fx['mu'][0]) * gain.cpu().numpy()[:, None, None]
    got_yhat = yhat.interior().permute(2, 0, 1).cpu().numpy()
    np.testing.assert_allclose(got_yhat, ref_yhat, rtol=1e-6, atol=1e-6)
    # replicate border of the bordered output
    full = yhat.t.view(yhat.rows, yhat.pitch, c)
    assert torch.equal(full[0, 1:1 + w], full[1, 1:1 + w]) and torch.equal(full[1:1 + h, 0], full[1:1 + h, 1])
    assert torch.equal(full[h + 1, 1:1 + w], full[h, 1:1 + w]) and torch.equal(full[0, 0], full[1, 1])
    # decoder side: scales -> host range decoder round trip through the real bitstream framing
    from aivc_b200 import entropy
    b = torch.empty(c * h * w, dtype=torch.float32, device=dev)
    _lib.check(L.aivc_laplace_scale(C.byref(fh), c, b.data_ptr(), _lib.stream_ptr()))
    sec = entropy.encode_y(got.reshape(c, h * w), nz.cpu().numpy())
    dec = entropy.decode_y(sec[4:], b.cpu().numpy().reshape(c, h, w), c, h, w)
    assert np.array_equal(dec, ref_q)


def _system(golden_dir, dev, precision):
    from aivc_b200 import models, gop as G
    from aivc_b200.codec import FrameCodec, planes_to_device
    from aivc_b200.plan import Config
    fx = np.load(os.path.join(golden_dir, 'system_80x112.npz'))
    h, w = int(fx['H']), int(fx['W'])
    net = models.build_standin(seed=4321, C=32, Cy=16, Cz=16, Csc=16)
    gop = G.generate_gop_struct('1_GOP_2')
    frames = {'frame_%d' % t: planes_to_device([fx['src_frame_%d_%s' % (t, k)] for k in 'yuv'], dev)
              for t in range(3)}
    codec = FrameCodec(net, h, w, dev, Config(precision=precision))
    bts, rec = codec.encode_gop(frames, gop)
    dec = codec.decode_gop(bts, gop)
    return fx, gop, bts, rec, dec


def test_codec_fp32_bit_exact_vs_oracle(golden_dir, dev):
    """fp32 engine: same bitstream bytes and same 8-bit reconstruction as the oracle, and the
    decoder reproduces the encoder's reconstruction (closed loop, like flag_bitstream_debug)."""
    fx, gop, bts, rec, dec = _system(golden_dir, dev, 'fp32')
    for f in gop:
        for a, b in zip(rec[f], dec[f]):
            assert torch.equal(a, b)
        for k, p in zip('yuv', rec[f]):
            assert np.array_equal(p.cpu().numpy(), fx['spec_rec_%s_%s' % (f, k)].reshape(-1)), (f, k)
        assert bts[f] == fx['spec_bytes_%s' % f].tobytes(), f


def test_oracle_decodes_gpu_bitstream(golden_dir, dev):
    """Cross check in the other direction: the oracle's CPU decoder reads the GPU encoder's
    bitstream and lands on the GPU encoder's reconstruction."""
    from aivc_b200 import models
    from oracle import codec_ref as O
    fx, gop, bts, rec, _ = _system(golden_dir, dev, 'fp32')
    net = models.build_standin(seed=4321, C=32, Cy=16, Cz=16, Csc=16)
    orec = O.decode_gop(net, O.Tables(net), bts, gop, int(fx['H']), int(fx['W']), cdf_mode='spec')
    for f in gop:
        for k, p in zip('yuv', rec[f]):
            got = np.rint(orec[f][k].numpy() * 255).astype(np.uint8).reshape(-1)
            assert np.array_equal(got, p.cpu().numpy()), (f, k)


@pytest.mark.parametrize('precision', ['fp32', 'bf16x3', 'bf16'])
@pytest.mark.parametrize('size', [(135, 241), (240, 416)])
def test_closed_loop_odd_and_reference_sizes(size, precision, dev):
    """Edge geometry: odd luma/chroma sizes, ceil-halving latents. Property:
    decode(encode(x)) reproduces the encoder's reconstruction exactly, for I, P and B."""
    from aivc_b200 import models, gop as G
    from aivc_b200.codec import FrameCodec, planes_to_device
    from aivc_b200.plan import Config
    h, w = size
    rng = np.random.default_rng(1)
    net = models.build_standin(seed=7, C=32, Cy=16, Cz=16, Csc=16)
    gop = G.generate_gop_struct('1_GOP_2')
    frames = {}
    for t in range(3):
        y = rng.integers(0, 256, (h, w), dtype=np.uint8)
        u = rng.integers(0, 256, ((h + 1) // 2, (w + 1) // 2), dtype=np.uint8)
        frames['frame_%d' % t] = planes_to_device([y, u, 255 - u], dev)
    codec = FrameCodec(net, h, w, dev, Config(precision=precision))
    bts, rec = codec.encode_gop(frames, gop)
    dec = codec.decode_gop(bts, gop)
    for f in gop:
        assert len(bts[f]) > 16
        for a, b in zip(rec[f], dec[f]):
            assert torch.equal(a, b)


def test_entropy_model_classes_vs_reference_golden(golden_dir, dev):
    """ParametricPdf / EntropyCoder / PdfParamParameterizer (mixture modes) / BallePdfEstim.forward mirrors against
    fixtures the REFERENCE's classes produced (oracle/gen_golden_pdf.py): pdf_estimator.py:17-70, 172-202,
    entropy_coder.py:18-30, misc_layers.py:172-269.  Tolerance: torch's expm1 / erf against ours, 2e-6 absolute on
    probabilities (values <= 1)."""
    import aivc_b200.layers as M
    fx = np.load(os.path.join(golden_dir, 'pdf_classes.npz'))
    C_ = 6
    for mode in ('laplace', 'laplace_two', 'normal_three_gamma'):
        x = torch.from_numpy(fx['pp_%s_x' % mode]).to(dev)
        prm = M.PdfParamParameterizer(mode, C_)(x)
        K = 1 + ('two' in mode) + 2 * ('three' in mode)
        assert len(prm) == K
        for k, d in enumerate(prm):
            assert np.array_equal(d['mu'].cpu().numpy(), fx['pp_%s_%d_mu' % (mode, k)])
            np.testing.assert_allclose(d['sigma'].cpu().numpy(), fx['pp_%s_%d_sigma' % (mode, k)], rtol=3e-7)
            np.testing.assert_allclose(d['gamma'].cpu().numpy(), fx['pp_%s_%d_gamma' % (mode, k)], rtol=3e-7)
            np.testing.assert_allclose(d['weight'].cpu().numpy(), fx['pp_%s_%d_weight' % (mode, k)], rtol=1e-6, atol=1e-7)
        y = torch.from_numpy(fx['pdf_%s_y' % mode]).to(dev)
        fam = 'normal' if 'normal' in mode else 'laplace'
        for zero_mu in (False, True):
            p = M.ParametricPdf(fam)(y, prm, zero_mu=zero_mu)
            np.testing.assert_allclose(p.cpu().numpy(), fx['pdf_%s_zero%d_p' % (mode, zero_mu)], rtol=0, atol=2e-6)
            rate = M.EntropyCoder()(torch.from_numpy(fx['pdf_%s_zero%d_p' % (mode, zero_mu)]).to(dev), y)
            np.testing.assert_allclose(rate.cpu().numpy(), fx['pdf_%s_zero%d_rate' % (mode, zero_mu)], rtol=1e-6, atol=1e-6)
    bz = M.BallePdfEstim(C_, pdf_family='')
    bz.load_state_dict({k[9:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith('balle_sd:')})
    with torch.no_grad():
        p = bz(torch.from_numpy(fx['balle_z']))
    np.testing.assert_allclose(p.numpy(), fx['balle_p'], rtol=1e-6, atol=1e-7)


def test_plan_cache_follows_the_weights_and_stays_out_of_the_module(golden_dir, dev, tmp_path):
    """(1) weights changed after the first forward (load_state_dict, fine-tuning) must not be ignored by the
    cached plan; (2) after a forward the module is still picklable / deep-copyable (the reference stores whole-module
    pickles) -- the caches live in a WeakKeyDictionary, not in module.__dict__."""
    import copy
    import io
    from aivc_b200 import plan
    m, fx = load_leaf('cheng_plain', golden_dir)
    x = torch.from_numpy(fx['x0']).to(dev)
    y0 = plan.run_module(m, x, _fp32()).cpu().numpy()
    np.testing.assert_allclose(y0, fx['y0'], rtol=RTOL, atol=ATOL)
    assert not any(k.startswith('_aivc') for k in m.__dict__)
    m2 = copy.deepcopy(m)
    buf = io.BytesIO()
    torch.save(m, buf)
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(0.5)
    y1 = plan.run_module(m, x, _fp32()).cpu().numpy()
    assert np.abs(y1 - y0).max() > 1e-3                      # the new weights are in use
    y2 = plan.run_module(m2, x, _fp32()).cpu().numpy()       # the copy kept the old ones
    assert np.array_equal(y2, y0)
    m.load_state_dict(m2.state_dict())
    assert np.array_equal(plan.run_module(m, x, _fp32()).cpu().numpy(), y0)
    assert len(plan.cached_plans(m)) == 1


@pytest.mark.parametrize('precision', ['fp32', 'bf16x3', 'bf16'])
def test_edge_inputs_empty_latents_and_tiny_frames(precision, dev):
    """Edge cases of the path (the reference handles them in bitstream.py:241-255, 292-296 and decode.py:562-571):
    (1) a flat frame whose latents quantise to ALL ZEROS in some nets -- no channel is sent, the y section is the single
        byte n_ch = 0 and the decoder must rebuild zeros without touching the range coder;
    (2) the smallest frames the four + two stride-2 stages allow without degenerate maps (18x34: y 2x3, z 1x1) and an odd
        size in both dimensions (33x47);
    (3) an all-intra GOP ('1_GOP_0') and a P-only chain ('LDP_3').
    Property everywhere: decoder == encoder reconstruction, and re-encoding gives the same bytes."""
    from aivc_b200 import models, gop as G, entropy
    from aivc_b200.codec import FrameCodec, planes_to_device
    from aivc_b200.plan import Config
    net = models.build_standin(seed=11, C=32, Cy=16, Cz=16, Csc=16)
    rng = np.random.default_rng(2)
    for (h, w), gop_name in (((18, 34), '1_GOP_2'), ((33, 47), 'LDP_3'), ((48, 64), '1_GOP_0')):
        gop = G.generate_gop_struct(gop_name)
        hc, wc = (h + 1) // 2, (w + 1) // 2
        frames = {}
        for i, f in enumerate(sorted(gop)):
            if i == 0:                                   # flat mid-grey frame
                frames[f] = planes_to_device([np.full((h, w), 128, np.uint8), np.full((hc, wc), 128, np.uint8),
                                              np.full((hc, wc), 128, np.uint8)], dev)
            else:
                frames[f] = planes_to_device([rng.integers(0, 256, (h, w), dtype=np.uint8),
                                              rng.integers(0, 256, (hc, wc), dtype=np.uint8),
                                              rng.integers(0, 256, (hc, wc), dtype=np.uint8)], dev)
        codec = FrameCodec(net, h, w, dev, Config(precision=precision))
        bts, rec = codec.encode_gop(frames, gop)
        dec = codec.decode_gop(bts, gop)
        bts2, _ = codec.encode_gop(frames, gop)
        assert bts2 == bts
        for f in gop:
            secs = entropy.split_sections(bts[f])
            assert len(secs) == 4
            for a, b in zip(rec[f], dec[f]):
                assert torch.equal(a, b), (h, w, gop_name, f)
    # (1) explicitly: a latent with no non-zero channel is one byte, and decodes to zeros
    sec = entropy.encode_y(np.zeros((4, 6), np.uint32), np.zeros(4, np.int32))
    assert sec == (1).to_bytes(4, 'big') + b'\x00'
    q = entropy.decode_y(sec[4:], np.ones((4, 2, 3), np.float32), 4, 2, 3)
    assert q.shape == (4, 2, 3) and not q.any()


@pytest.mark.parametrize('size', [(80, 112), (135, 241)])
def test_reference_style_decoder_on_the_dropin_modules(size, golden_dir, dev):
    """The reference's decoder composition (real_life/decode.py:455-898, restated call for call in
    tests/module_decoder.py) driven through the drop-in modules' own forward() on CUDA tensors -- the path a user gets
    who lets the reference's Decoder run on the mirrors -- lands on the fused FrameCodec's reconstruction: exact fp32
    engine on both sides, so the planes are identical (and, at 80x112, equal to the oracle's golden reconstruction)."""
    from aivc_b200 import models, gop as G, plan
    from aivc_b200.codec import FrameCodec, planes_to_device
    from aivc_b200.plan import Config
    from tests import module_decoder
    h, w = size
    net = models.build_standin(seed=4321, C=32, Cy=16, Cz=16, Csc=16).to(dev)
    gop = G.generate_gop_struct('1_GOP_2')
    if size == (80, 112):
        fx = np.load(os.path.join(golden_dir, 'system_80x112.npz'))
        clip = [[fx['src_frame_%d_%s' % (t, k)] for k in 'yuv'] for t in range(3)]
    else:
        rng = np.random.default_rng(8)
        clip = [[rng.integers(0, 256, s, dtype=np.uint8) for s in ((h, w), ((h + 1) // 2, (w + 1) // 2), ((h + 1) // 2, (w + 1) // 2))]
                for _ in range(3)]
    frames = {'frame_%d' % t: planes_to_device(clip[t], dev) for t in range(3)}
    cfg = Config(precision='fp32')
    bts, rec = FrameCodec(net, h, w, dev, cfg).encode_gop(frames, gop)
    old = plan.set_default_config(cfg)
    try:
        dec = module_decoder.decode_gop(net, bts, gop, h, w, dev)
    finally:
        plan.set_default_config(old)
    for f in gop:
        for k, p in zip('yuv', rec[f]):
            got = torch.round(dec[f][k] * 255).to(torch.uint8).reshape(-1)
            assert torch.equal(got, p), (f, k, int((got.int() - p.int()).abs().max()))
            if size == (80, 112):
                assert np.array_equal(got.cpu().numpy(), fx['spec_rec_%s_%s' % (f, k)].reshape(-1))
