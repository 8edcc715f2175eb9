"""CPU suite: the oracle against the committed golden vectors (which gen_golden.py pinned to
the reference classes), without /root/reference."""
import os

import numpy as np
import pytest
import torch

from oracle import nn_ref as R, codec_ref as C
from tests.leafcfg import LEAVES, load_leaf


@pytest.mark.parametrize('name', sorted(LEAVES))
def test_oracle_reproduces_leaf_golden(name, golden_dir):
    m, fx = load_leaf(name, golden_dir)
    with torch.no_grad():
        for i in range(2):
            y = R.forward_module(m, torch.from_numpy(fx['x%d' % i]))
            assert torch.equal(y, torch.from_numpy(fx['y%d' % i]))


def test_oracle_pixel_ends(golden_dir):
    fx = np.load(os.path.join(golden_dir, 'misc.npz'))
    for tag in ('even', 'odd'):
        yuv = {k: torch.from_numpy(fx['%s_in_%s' % (tag, k)]) for k in 'yuv'}
        x444 = R.input_layer(yuv)
        assert torch.equal(x444, torch.from_numpy(fx[tag + '_x444']))
        h, w = x444.shape[2:]
        fin = R.finalize_frame(x444 * 1.3 - 0.1, h, w)
        for k in 'yuv':
            assert torch.equal(fin[k], torch.from_numpy(fx['%s_fin_%s' % (tag, k)]))
        wr = R.warp(x444, torch.from_numpy(fx[tag + '_flow']))
        assert torch.equal(wr, torch.from_numpy(fx[tag + '_warp']))
    mu, sigma = R.mu_sigma(torch.from_numpy(fx['hs_out']), 8)
    assert torch.equal(mu, torch.from_numpy(fx['mu'])) and torch.equal(sigma, torch.from_numpy(fx['sigma']))


def test_laplace_spec_matches_torch_reference_table(golden_dir):
    """The deterministic integer Laplace CDF vs the table the reference builds with torch fp32."""
    fx = np.load(os.path.join(golden_dir, 'misc.npz'))
    spec = C.laplace_table_spec(fx['laplace_sigma'])
    ref = fx['laplace_ref_u16']
    d = np.abs(spec.astype(np.int32) - ref.astype(np.int32))
    assert d.max() <= 1
    assert (d != 0).mean() < 1e-3
    # strictly increasing rows, as the range coder needs (entry 513 wraps to 0 and is never
    # read: the coder substitutes 0x10000 for the top symbol)
    assert (np.diff(spec[:, :513].astype(np.int32), axis=1) > 0).all()


def test_system_golden_decodes(golden_dir):
    """Oracle decoder on the committed bitstreams reproduces the committed reconstruction."""
    from aivc_b200 import models, gop as G
    fx = np.load(os.path.join(golden_dir, 'system_80x112.npz'))
    h, w = int(fx['H']), int(fx['W'])
    net = models.build_standin(seed=4321, C=32, Cy=16, Cz=16, Csc=16)
    tables = C.Tables(net)
    gop = G.generate_gop_struct('1_GOP_2')
    for mode in ('spec', 'reference'):
        bts = {f: fx['%s_bytes_%s' % (mode, f)].tobytes() for f in gop}
        rec = C.decode_gop(net, tables, bts, gop, h, w, cdf_mode=mode)
        for f in gop:
            for k in 'yuv':
                got = (rec[f][k].numpy() * 255).round().astype(np.uint8)
                assert np.array_equal(got, fx['%s_rec_%s_%s' % (mode, f, k)])


def test_metrics_oracle_reproduces_reference_golden(golden_dir):
    """oracle/metrics_ref.py against the values the reference's own MSELoss / MSSSIMLoss gave for the same
    seeded planes (oracle/gen_golden_metrics.py): odd sizes, windows shrinking below 11 at the small scales."""
    import os
    from oracle import metrics_ref as M, gen_golden_metrics as Gm
    cases = np.load(os.path.join(golden_dir, 'metrics.npz'))['cases']
    assert len(cases) == len(Gm.CASES)
    for seed, h, w, mse, ms in cases:
        a, b = Gm.planes(int(seed), int(h), int(w))
        got = M.frame_metrics(Gm.as_dic(a), Gm.as_dic(b))
        assert abs(got['mse'] - mse) <= 1e-9 and abs(got['ms_ssim'] - ms) <= 2e-6


def test_reference_decoder_reads_the_cuda_encoders_golden_bitstream():
    """The REFERENCE's own real_life.decode.Decoder + ArithmeticCoder (torchac shimmed, transforms evaluated by the
    oracle) decode the golden GOP's 'spec' bitstream -- byte for byte what the CUDA bf16x3 and fp32 encoders write
    (tests/test_gpu_engine_x3.py::test_codec_x3_bytes_identical_to_oracle, test_gpu_parity.py::
    test_codec_fp32_bit_exact_vs_oracle) -- to exactly the encoder's reconstruction.  Needs the reference tree."""
    import os
    import numpy as np
    import pytest
    if not os.path.isdir('/root/reference/src'):
        pytest.skip('reference tree not present')
    from aivc_b200 import models, gop as G
    from oracle.ref_decoder import build_reference_decoder, reference_decode_gop
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'system_80x112.npz'))
    gop = G.generate_gop_struct('1_GOP_2')
    net = models.build_standin(seed=4321, C=32, Cy=16, Cz=16, Csc=16)
    dec = reference_decode_gop(build_reference_decoder(net), {f: fx['spec_bytes_%s' % f].tobytes() for f in gop}, gop,
                               int(fx['H']), int(fx['W']))
    for f in gop:
        for k in 'yuv':
            got = np.rint(dec[f][k].numpy() * 255).astype(np.uint8).reshape(-1)
            assert np.array_equal(got, fx['spec_rec_%s_%s' % (f, k)].reshape(-1)), (f, k)


def test_synthetic_sources_of_the_config_fixtures_are_reproducible():
    """tests/synth.py is integer-only, so the frames the BASELINE-configuration fixtures were minted from must come out
    bit-identical on every machine and numpy build (a change here would make the GPU parity tests compare against
    fixtures of different pictures)."""
    import hashlib
    from tests import synth
    want = {(720, 3, 720, 1280): '58032c58517671812a3d80d01cd7dc46',
            (1080, 3, 1080, 1920): '9751ef096fb0d2f6d111de30ed03a7ca',
            (1081, 9, 1080, 1920): '7d9da5401645eacc704505c58583e209'}
    for (seed, n, h, w), md5 in want.items():
        clip = synth.clip(seed, n, h, w)
        assert clip[0][0].shape == (h, w) and clip[0][1].shape == (h // 2, w // 2) and clip[0][0].dtype.name == 'uint8'
        assert hashlib.md5(b''.join(p.tobytes() for fr in clip for p in fr)).hexdigest() == md5, (seed, n, h, w)


def test_config_fixture_is_what_the_oracle_produces_here():
    """The smallest BASELINE-configuration fixture (416x240 real frame, all intra, C=128) re-minted by the oracle on this
    machine: same latent indices, same bytes, same planes as the committed tests/golden/cfg_bubbles240.npz."""
    import hashlib
    import os
    import numpy as np
    import torch
    from aivc_b200 import models
    from oracle import codec_ref as O
    from tests import parity_cfg
    fx = np.load(os.path.join(parity_cfg.GOLDEN, 'cfg_bubbles240.npz'))
    y, u, v = parity_cfg.source_frames('bubbles240', fx)[0]
    yuv = {k: torch.from_numpy(p.astype(np.float32) / 255.)[None, None] for k, p in zip('yuv', (y, u, v))}
    net = models.build_standin(**parity_cfg.MODELS['bubbles240'])
    z = O.zero_yuv(240, 416)
    data, rec, aux = O.encode_frame(net, O.Tables(net), yuv, z, z, 0)
    assert np.array_equal(aux['codec']['q'].numpy().astype(np.int32)[0], fx['frame_0_codec_q'].astype(np.int32))
    assert np.array_equal(aux['codec']['z_hat'].numpy().astype(np.int32)[0], fx['frame_0_codec_z'].astype(np.int32))
    assert hashlib.md5(data).hexdigest() == str(fx['frame_0_bytes_md5']) and len(data) == int(fx['frame_0_nbytes'])
    planes = [np.rint(rec[k].numpy() * 255).astype(np.uint8)[0, 0] for k in 'yuv']
    assert hashlib.md5(b''.join(p.tobytes() for p in planes)).hexdigest() == str(fx['frame_0_planes_md5'])
