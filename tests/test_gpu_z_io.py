"""GPU suite, part 3 (runs last: an IO / adapter failure must not hide the kernel parity tests): the
reference-facing entry points (GOP_forward contract, video container round trip), the direct .yuv file path,
multi-GOP decoding with the GPU lagging behind the host, run-to-run determinism, per-frame metrics."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU suite needs a CUDA device'
    return torch.device('cuda:0')


def test_gop_forward_contract_and_video_roundtrip(golden_dir, dev, tmp_path):
    """FullNet.GOP_forward with the reference's model_input dict (model_management.py:307-317):
    net_out keys, the GOP file it leaves behind, and decode_video of the assembled .bin."""
    from aivc_b200 import models, gop as G, container, adapter
    from aivc_b200.plan import Config
    fx = np.load(os.path.join(golden_dir, 'system_80x112.npz'))
    h, w = int(fx['H']), int(fx['W'])
    net = models.build_standin(seed=4321, C=32, Cy=16, Cz=16, Csc=16)
    gop = G.generate_gop_struct('1_GOP_2')
    raw = {}
    for t in range(3):
        raw['frame_%d' % t] = {}
        for k in 'yuv':
            a = fx['src_frame_%d_%s' % (t, k)].astype(np.float32) / 255.
            raw['frame_%d' % t][k] = torch.from_numpy(a).reshape(1, 1, *a.shape[-2:]).to(dev)
    d = str(tmp_path) + '/bs/'
    model_input = {'GOP_struct': gop, 'GOP_struct_name': '1_GOP_2', 'raw_frames': raw, 'idx_rate': 0.,
                   'index_GOP_in_video': 0, 'generate_bitstream': True, 'real_idx_first_frame': 0,
                   'bitstream_dir': d, 'flag_bitstream_debug': False}
    out = adapter.gop_forward(net, model_input, cfg=Config(precision='fp32'))
    for f in gop:
        for key in ('x_hat', 'alpha', 'beta', 'warping', 'code', 'mode_rate_y', 'mode_rate_z', 'codec_rate_y',
                    'codec_rate_z'):
            assert key in out[f], key
        for k in 'yuv':
            got = (out[f]['x_hat'][k].cpu().numpy() * 255).round().astype(np.uint8).reshape(-1)
            assert np.array_equal(got, fx['spec_rec_%s_%s' % (f, k)].reshape(-1))     # fp32 engine == oracle
    # the logging tensors are the real ones: rate estimates, alpha, beta, warping against the
    # oracle's evaluation of the same frames (fp32 engine: same symbols, so the only difference is fp32 rounding)
    from oracle import codec_ref as O, nn_ref as R
    tables = O.Tables(net)
    yuv = {f: {k: raw[f][k].cpu() for k in 'yuv'} for f in raw}
    orec = {}
    for f in sorted(gop, key=lambda f: gop[f]['coding_order']):
        t = gop[f]['type']
        prev = orec[gop[f]['prev_ref']] if t != 0 else O.zero_yuv(h, w)
        nxt = orec[gop[f]['next_ref']] if t == 2 else O.zero_yuv(h, w)
        _, orec[f], aux = O.encode_frame(net, tables, yuv[f], prev, nxt, t)
        o = out[f]
        for net_name, key, cn in (('mof', 'mode', net.mode_net.mode_net), ('codec', 'codec', net.codec_net.codec_net)):
            if net_name not in aux:
                assert float(o[key + '_rate_y'].sum()) == 0.0 and float(o[key + '_rate_z'].sum()) == 0.0
                continue
            ref_y = R.laplace_rate_bits(aux[net_name]['q'], aux[net_name]['sigma'])
            np.testing.assert_allclose(o[key + '_rate_y'].cpu().numpy(), ref_y.numpy(), rtol=2e-5, atol=2e-5)
            with torch.no_grad():
                ref_z = -torch.log2(torch.clamp(cn.pdf_z(aux[net_name]['z_hat']), 2.0 ** -16, 1.0))
            np.testing.assert_allclose(o[key + '_rate_z'].cpu().numpy(), ref_z.numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(o['alpha'].cpu().numpy(), aux['alpha'].expand(1, 3, h, w).numpy(), atol=2e-5)
        np.testing.assert_allclose(o['warping'].cpu().numpy(), aux['x_warp'].numpy(), atol=2e-5)
        assert o['beta'].shape == (1, 3, h, w) and float(o['beta'].min()) >= 0.0 and float(o['beta'].max()) <= 1.0
        if t == 1:
            assert float(o['beta'].min()) == 1.0                       # P frames: beta = 1 (decode.py:736-739)
        # estimated rate vs real coded size: the range coder is within a few % + a few bytes of the estimate
        est = float(o['codec_rate_y'].sum() + o['codec_rate_z'].sum())
        real = o['coded_bits']['codec_y'] + o['coded_bits']['codec_z']
        assert abs(real - est) <= 0.1 * est + 256, (est, real)
    m = adapter.compute_metrics_one_gop(out, raw)
    assert set(m) == set(gop) | {'GOP'} and m['GOP']['total_rate_bpp'] > 0 and 0 < m['frame_0']['ms_ssim'] <= 1
    gbytes = open(d + '0g', 'rb').read()
    name, rate, frames = container.unpack_gop(gbytes)
    assert name == '1_GOP_2' and [bytes(b) for b in frames] == [fx['spec_bytes_frame_%d' % i].tobytes() for i in range(3)]
    # the same content through the video container and back
    from aivc_b200.codec import latent_dims
    dy, dz = latent_dims(h, w)
    video = container.pack_video((h, w), dy, dz, [gbytes], 0, 2)
    dec, dims, first, last = adapter.decode_video(net, video, dev, Config(precision='fp32'))
    assert dims['x'] == (h, w) and (first, last) == (0, 2)
    for f in gop:
        for k, p in zip('yuv', dec[0][f]):
            assert np.array_equal(p.cpu().numpy(), fx['spec_rec_%s_%s' % (f, k)].reshape(-1))


def test_yuv_file_to_bitstream_to_yuv_file(dev, tmp_path):
    """Direct .yuv path: 4 frames, GOP of 3 -> two GOPs, the second padded; the
    decoded file holds exactly the encoder's reconstructions of the 4 real frames."""
    from aivc_b200 import models, adapter, yuvio, gop as G
    from aivc_b200.plan import Config
    w, h = 64, 48
    rng = np.random.default_rng(9)
    path = str(tmp_path / ('clip_%dx%d_25_420.yuv' % (w, h)))
    with yuvio.YuvWriter(path) as wr:
        for _ in range(4):
            wr.append((rng.integers(0, 256, w * h, dtype=np.uint8), rng.integers(0, 256, w * h // 4, dtype=np.uint8),
                       rng.integers(0, 256, w * h // 4, dtype=np.uint8)))
    net = models.build_standin(seed=3, C=32, Cy=16, Cz=16, Csc=16)
    cfg = Config(precision='fp32')
    video = adapter.encode_yuv(net, path, '1_GOP_2', device=dev, cfg=cfg)
    out = str(tmp_path / 'dec_64x48_25_420.yuv')
    assert adapter.decode_to_yuv(net, video, out, device=dev, cfg=cfg) == 4
    # reference: the same GOPs through encode_gop directly
    rd, dec = yuvio.YuvReader(path), yuvio.YuvReader(out)
    assert len(dec) == 4
    codec = adapter.codec_for(net, h, w, dev, cfg)
    gop = G.generate_gop_struct('1_GOP_2')
    k = 0
    for first, keep in yuvio.gop_schedule(0, 3, 3):
        _, rec = codec.encode_gop(rd.gop_frames(first, 3, 3, dev), gop)
        for j in range(keep):
            for a, b in zip(rec['frame_%d' % j], dec.frame(k)):
                assert np.array_equal(a.cpu().numpy(), b)
            k += 1


@pytest.mark.parametrize('case', [0, 1, 2])
def test_frame_metrics_vs_oracle(case, dev):
    """aivc_frame_metrics (MSE / PSNR / plane-weighted MS-SSIM on the device) against the oracle pinned to the
    reference's loss_function classes; tolerance: separable fp32 filtering vs direct 2-D convolution."""
    from aivc_b200 import metrics
    from oracle import metrics_ref as M, gen_golden_metrics as Gm
    seed, h, w = Gm.CASES[case]
    a, b = Gm.planes(seed, h, w)
    ref = M.frame_metrics(Gm.as_dic(a), Gm.as_dic(b))
    to_dev = lambda pl: tuple(torch.from_numpy(np.ascontiguousarray(p).reshape(-1)).to(dev) for p in pl)
    got = metrics.frame_metrics(to_dev(a), to_dev(b), h, w)
    assert abs(got['mse'] - ref['mse']) <= 1e-6 * ref['mse'] + 1e-12
    assert abs(got['psnr'] - ref['psnr']) <= 1e-4
    assert abs(got['ms_ssim'] - ref['ms_ssim']) <= 2e-5
    again = metrics.frame_metrics(to_dev(a), to_dev(b), h, w)
    assert again == got                                    # fixed-order reductions: run-to-run identical


def _noise_gops(n_gops, h, w, seed, dev):
    from aivc_b200.codec import planes_to_device
    rng = np.random.default_rng(seed)
    hc, wc = (h + 1) // 2, (w + 1) // 2
    return [{'frame_%d' % t: planes_to_device([rng.integers(0, 256, (h, w), dtype=np.uint8),
                                               rng.integers(0, 256, (hc, wc), dtype=np.uint8),
                                               rng.integers(0, 256, (hc, wc), dtype=np.uint8)], dev)
             for t in range(3)} for _ in range(n_gops)]


@pytest.mark.parametrize('precision', ['fp32', 'bf16x3', 'bf16'])
def test_multi_gop_decode_with_lagging_gpu(precision, dev):
    """Round-1 driver failure, reproduced on purpose: decode_video enqueues GOP after GOP without a stream
    synchronisation, so the host runs ahead of the GPU.  Here the GPU is held back by a long sleep kernel and
    the GOPs differ, so any host write into staging memory that a queued copy has not read yet (the pinned z
    slot re-read by the synthesis pass was one) shows up as a decoder / encoder mismatch."""
    from aivc_b200 import models, adapter, container, gop as G
    from aivc_b200.codec import latent_dims
    from aivc_b200.plan import Config
    h, w = 64, 48
    net = models.build_standin(seed=3, C=32, Cy=16, Cz=16, Csc=16)
    cfg = Config(precision=precision)
    gop = G.generate_gop_struct('1_GOP_2')
    order = sorted(gop, key=lambda f: int(f.split('_')[1]))
    codec = adapter.codec_for(net, h, w, dev, cfg)
    gops = _noise_gops(4, h, w, 21, dev)
    packed, recs = [], []
    for frames in gops:
        bts, rec = codec.encode_gop(frames, gop)
        packed.append(container.pack_gop('1_GOP_2', [bts[f] for f in order], 0.))
        recs.append({f: tuple(p.clone() for p in rec[f]) for f in rec})
    dy, dz = latent_dims(h, w)
    video = container.pack_video((h, w), dy, dz, packed, 0, 3 * len(gops) - 1)
    # hold the GPU back in front of every synthesis pass: reconstruction of GOP g is still queued when the host
    # starts on GOP g + 1
    engines = [e for c in codec._lanes() for e in (c.mof, c.codec)]         # (every lane of frames in flight)
    for eng in engines:
        orig = eng.synth_launch
        def slow(*a, _orig=orig, **k):
            torch.cuda._sleep(20_000_000)               # ~10 ms
            return _orig(*a, **k)
        eng.synth_launch = slow
    try:
        for _ in range(2):
            dec, _, _, _ = adapter.decode_video(net, video, dev, cfg)
            for g, rec in enumerate(recs):
                for f in order:
                    for a, b in zip(rec[f], dec[g][f]):
                        assert torch.equal(a, b), (g, f)
    finally:
        for eng in engines:
            del eng.synth_launch


@pytest.mark.parametrize('precision', ['fp32', 'bf16x3', 'bf16'])
def test_encoder_is_run_to_run_deterministic(precision, dev):
    """The same GOP coded 20 times on one codec (and once on a fresh one) gives identical bytes and identical
    planes: fixed accumulation order, no atomics, no stale staging memory.  This is what
    src/sanity_script.sh:3 / func_util/cluster_mngt.py:27-37 protect in the reference."""
    from aivc_b200 import models, gop as G
    from aivc_b200.codec import FrameCodec
    from aivc_b200.plan import Config
    h, w = 80, 112
    net = models.build_standin(seed=4321, C=32, Cy=16, Cz=16, Csc=16)
    gop = G.generate_gop_struct('1_GOP_2')
    a, b = _noise_gops(2, h, w, 5, dev)
    codec = FrameCodec(net, h, w, dev, Config(precision=precision))
    bts0, rec0 = codec.encode_gop(a, gop)
    rec0 = {f: tuple(p.clone() for p in rec0[f]) for f in rec0}
    for i in range(20):
        if i % 3 == 1:
            codec.encode_gop(b, gop)                    # other content in between: no state may leak
        if i % 5 == 2:
            torch.cuda._sleep(50_000_000)               # vary the host / device skew
        bts, rec = codec.encode_gop(a, gop)
        assert bts == bts0, i
        for f in rec0:
            for p, q in zip(rec0[f], rec[f]):
                assert torch.equal(p, q), (i, f)
    fresh = FrameCodec(net, h, w, dev, Config(precision=precision))
    bts, rec = fresh.encode_gop(a, gop)
    assert bts == bts0
    dec = fresh.decode_gop(bts, gop)
    for f in rec0:
        for p, q, r in zip(rec0[f], rec[f], dec[f]):
            assert torch.equal(p, q) and torch.equal(p, r), f


@pytest.mark.parametrize('precision', ['bf16x3', 'bf16'])
def test_two_frames_in_flight_same_bitstream(precision, dev):
    """Config.frames_in_flight = 2 deals the frames of a dependency level to two streams (two sets of plans and
    buffers, events on the reconstructed planes).  Every frame is still coded by the same kernels in the same order:
    bytes and planes are identical to the one-lane codec's, and the two-lane decoder reproduces them."""
    from aivc_b200 import models, gop as G
    from aivc_b200.codec import FrameCodec
    from aivc_b200.plan import Config
    h, w = 80, 112
    net = models.build_standin(seed=4321, C=32, Cy=16, Cz=16, Csc=16)
    gop = G.generate_gop_struct('1_GOP_8')               # levels of 1, 1, 1, 2, 4 frames
    rng = np.random.default_rng(17)
    from aivc_b200.codec import planes_to_device
    frames = {f: planes_to_device([rng.integers(0, 256, (h, w), dtype=np.uint8), rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8),
                                   rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)], dev) for f in gop}
    one = FrameCodec(net, h, w, dev, Config(precision=precision, frames_in_flight=1))
    two = FrameCodec(net, h, w, dev, Config(precision=precision, frames_in_flight=2))
    b1, r1 = one.encode_gop(frames, gop)
    for rep in range(3):
        if rep == 1:
            torch.cuda._sleep(100_000_000)               # vary the skew between the lanes
        b2, r2 = two.encode_gop(frames, gop)
        d2 = two.decode_gop(b2, gop)
        assert b2 == b1
        for f in gop:
            for x, y, z in zip(r1[f], r2[f], d2[f]):
                assert torch.equal(x, y) and torch.equal(x, z), (rep, f)
    assert len(two._lanes()) == 2 and len(one._lanes()) == 1


@pytest.mark.parametrize('precision', ['bf16x3', 'bf16'])
def test_cuda_graph_replay_same_bitstream(precision, dev):
    """Config.cuda_graphs: every transform captured once (programmatic-dependent-launch edges, two-lane forks / joins
    of the attention blocks) and replayed with one cudaGraphLaunch.  Same kernels, same order: bytes and planes equal
    the plain-launch codec's over repeated GOPs, including the frame-type dependent gain pointer of g_a."""
    from aivc_b200 import models, gop as G
    from aivc_b200.codec import FrameCodec
    from aivc_b200.plan import Config
    h, w = 80, 112
    net = models.build_standin(seed=4321, C=32, Cy=16, Cz=16, Csc=16)
    gop = G.generate_gop_struct('1_GOP_4')
    a, b = _noise_gops(2, h, w, 23, dev)
    frames = [dict(a, frame_3=b['frame_0'], frame_4=b['frame_1']), dict(b, frame_3=a['frame_2'], frame_4=a['frame_0'])]
    plain = FrameCodec(net, h, w, dev, Config(precision=precision, cuda_graphs=False))
    graph = FrameCodec(net, h, w, dev, Config(precision=precision, cuda_graphs=True))
    for rep in range(3):                        # rep 0: warm-up runs, rep 1: capture, rep 2: replay
        for fr in frames:
            b1, r1 = plain.encode_gop(fr, gop)
            b2, r2 = graph.encode_gop(fr, gop)
            d2 = graph.decode_gop(b2, gop)
            assert b1 == b2, rep
            for f in gop:
                for x, y, z in zip(r1[f], r2[f], d2[f]):
                    assert torch.equal(x, y) and torch.equal(x, z), (rep, f)
    assert any(p._graphs for p in (graph.codec.g_a, graph.codec.g_s, graph.mof.g_a))


@pytest.mark.parametrize('precision', ['bf16x3', 'bf16', 'fp32'])
def test_float_planes_and_uint8_planes_agree(precision, dev):
    """encode_gop takes uint8 planes or fp32 planes in [0, 1].  Float planes holding 8-bit levels (what the reference
    feeds: PNG-sourced frames / 255) give the same bytes as the uint8 planes -- in the split-bf16 mode the stages that
    skip the lo halves of an exact input switch to the full product for float planes --, and float planes that are NOT
    8-bit levels still round-trip (decoder == encoder)."""
    from aivc_b200 import models, gop as G
    from aivc_b200.codec import FrameCodec
    from aivc_b200.plan import Config
    h, w = 80, 112
    net = models.build_standin(seed=4321, C=32, Cy=16, Cz=16, Csc=16)
    gop = G.generate_gop_struct('1_GOP_2')
    u8 = _noise_gops(1, h, w, 31, dev)[0]
    as_float = {f: tuple(p.float() / 255. for p in u8[f]) for f in u8}
    codec = FrameCodec(net, h, w, dev, Config(precision=precision))
    b_u8, r_u8 = codec.encode_gop(u8, gop)
    b_f, r_f = codec.encode_gop(as_float, gop)
    assert b_f == b_u8
    for f in gop:
        for x, y in zip(r_u8[f], r_f[f]):
            assert torch.equal(x, y)
    g = torch.Generator(device='cpu').manual_seed(1)
    odd = {f: tuple(torch.rand(p.shape, generator=g).to(dev) for p in u8[f]) for f in u8}     # not 8-bit levels
    b_o, r_o = codec.encode_gop(odd, gop)
    d_o = codec.decode_gop(b_o, gop)
    for f in gop:
        for x, y in zip(r_o[f], d_o[f]):
            assert torch.equal(x, y)
    b_again, _ = codec.encode_gop(u8, gop)                    # and back to uint8 planes: unchanged bytes
    assert b_again == b_u8


def test_closed_loop_2160p(dev):
    """Maximum size of practical interest (3840x2160, four times the benchmark's pixels): tensor-map extents, buffer
    offsets beyond 2^31 bytes and the persistent kernels' tile counters at 4K; I, P and B frame, default engine,
    small-width stand-in (the geometry is what is being tested)."""
    from aivc_b200 import models, gop as G
    from aivc_b200.codec import FrameCodec
    from aivc_b200.plan import Config
    h, w = 2160, 3840
    net = models.build_standin(seed=5, C=64, Cy=32, Cz=32, Csc=32)
    gop = G.generate_gop_struct('1_GOP_2')
    g = torch.Generator(device='cpu').manual_seed(4)
    frames = {f: tuple(torch.randint(0, 256, (n,), dtype=torch.uint8, generator=g).to(dev)
                       for n in (h * w, h * w // 4, h * w // 4)) for f in gop}
    codec = FrameCodec(net, h, w, dev, Config())
    bts, rec = codec.encode_gop(frames, gop)
    dec = codec.decode_gop(bts, gop)
    for f in gop:
        assert len(bts[f]) > 1000
        for a, b in zip(rec[f], dec[f]):
            assert torch.equal(a, b), f
    del codec
    torch.cuda.empty_cache()
