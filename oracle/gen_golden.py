"""ORACLE (test infrastructure only).  Run HERE (where /root/reference exists):

    python -m oracle.gen_golden

Pins ``oracle/nn_ref.py`` and ``oracle/codec_ref.py`` against the reference's own Python
classes imported from /root/reference/src (never copied), executed by the same torch build,
and writes small input/output fixtures to ``tests/golden/*.npz``.  The GPU box has no
/root/reference: tests there only read the fixtures.

What is checked while generating (any failure aborts):
  * every mirrored leaf class: reference(x) == oracle(reference module)(x) == oracle(mirror
    module with the reference's state_dict)(x), bit for bit;
  * warp / InputLayer / OutputLayer / PdfParamParameterizer / GainMatrix / BallePdfEstim.cdf /
    cast_before_png_saving / GOP structures / ArithmeticCoder's float CDF tables;
  * system level: a bitstream written by the oracle encoder (reference-CDF mode) is decoded by
    the reference's own ``real_life.decode.Decoder`` + ``ArithmeticCoder`` (torchac replaced by
    oracle/torchac_shim.py) to exactly the frames the oracle decoder produces.
"""
import hashlib
import io
import os
import sys
import tempfile
import contextlib

import numpy as np
import torch

REF_SRC = '/root/reference/src'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'tests', 'golden')


def _np(t):
    return t.detach().cpu().numpy()


def state_np(m):
    return {k: _np(v) for k, v in m.state_dict().items()}


def sd_hash(m):
    h = hashlib.sha256()
    for k, v in sorted(m.state_dict().items()):
        h.update(k.encode())
        h.update(_np(v).tobytes())
    return h.hexdigest()


def main():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF_SRC)
    from oracle import nn_ref as R, codec_ref as C, torchac_shim
    sys.modules['torchac'] = torchac_shim
    import aivc_b200.layers as M
    from aivc_b200 import models, gop as G

    with contextlib.redirect_stdout(io.StringIO()):
        from layers.misc import custom_conv_layers as rc, misc_layers as rm, attention as ra
        from layers.ae import ae_layers as rae
        from layers.multi_rate.gain_matrix import GainMatrix as RefGain
        from layers.entropy_coding.pdf_estimator import BallePdfEstim as RefBalle
        from func_util.optical_flow import warp as ref_warp
        from func_util.img_processing import cast_before_png_saving
        from func_util.GOP_structure import generate_gop_struct as ref_gop
        from real_life.bitstream import ArithmeticCoder
        from real_life.decode import Decoder as RefDecoder
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)
    torch.set_num_threads(4)

    # ------------------------------------------------------------- leaf classes
    def randomize_gdn(m, seed):
        g = torch.Generator().manual_seed(seed)
        for name, p in m.named_parameters():
            if name.endswith('gamma'):
                p.add_(0.05 * torch.rand(p.shape, generator=g))
            if name.endswith('beta'):
                p.add_(0.2 * torch.rand(p.shape, generator=g))

    C16 = 16
    leaves = [
        ('conv_k5_s2_gdn', lambda L: L.CustomConvLayer(5, 9, C16, non_linearity='gdn', conv_stride=2), 9),
        ('conv_k3_s1_leaky', lambda L: L.CustomConvLayer(3, C16, C16, non_linearity='leaky_relu'), C16),
        ('conv_k3_s2_no', lambda L: L.CustomConvLayer(3, C16, 8, non_linearity='no', conv_stride=2), C16),
        ('conv_k3_s1_relu', lambda L: L.CustomConvLayer(3, C16, C16, non_linearity='relu'), C16),
        ('conv_k3_s1_igdn', lambda L: L.CustomConvLayer(3, C16, C16, non_linearity='gdn_inverse'), C16),
        ('up_k3_leaky', lambda L: L.UpscalingLayer(3, C16, C16, non_linearity='leaky_relu'), C16),
        ('up_k5_no', lambda L: L.UpscalingLayer(5, C16, 3, non_linearity='no'), C16),
        ('up_k3_igdn', lambda L: L.UpscalingLayer(3, 8, C16, non_linearity='gdn_inverse'), 8),
        ('cheng_plain', lambda L: L.ChengResBlock(C16, 'plain'), C16),
        ('cheng_down', lambda L: L.ChengResBlock(C16, 'down'), C16),
        ('cheng_up', lambda L: L.ChengResBlock(C16, 'up_tconv'), C16),
        ('resblock', lambda L: L.ResBlock(3, C16), C16),
        ('attresblock', lambda L: L.AttentionResBlock(C16), C16),
        ('attention', lambda L: L.SimplifiedAttention(C16), C16),
        ('attention_light', lambda L: L.SimplifiedAttention(C16, lightweight_resblock=True), C16),
    ]

    class RefNS:       # the reference classes under one namespace
        pass
    for mod in (rc, rm, ra):
        for k in dir(mod):
            setattr(RefNS, k, getattr(mod, k))

    sizes = [(17, 23), (24, 32)]
    for li, (name, mk, cin) in enumerate(leaves):
        torch.manual_seed(100 + li)
        with contextlib.redirect_stdout(io.StringIO()):
            ref = mk(RefNS).eval()
        randomize_gdn(ref, 200 + li)
        mir = mk(M).eval()
        missing = mir.load_state_dict(ref.state_dict(), strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        rec = {'class': name}
        for si, (h, w) in enumerate(sizes):
            x = torch.randn(1, cin, h, w, generator=torch.Generator().manual_seed(300 + li + si))
            y_ref = ref(x)
            y_o1 = R.forward_module(ref, x)
            y_o2 = R.forward_module(mir, x)
            assert torch.equal(y_ref, y_o1), name + ': oracle != reference'
            assert torch.equal(y_ref, y_o2), name + ': oracle(mirror) != reference'
            rec['x%d' % si], rec['y%d' % si] = _np(x), _np(y_ref)
        for k, v in state_np(ref).items():
            rec['sd:' + k] = v
        np.savez_compressed(os.path.join(OUT, 'leaf_%s.npz' % name), **rec)
        print('leaf', name, 'ok', tuple(y_ref.shape))

    # ------------------------------------------------------------- pixel-end / entropy helpers
    g = torch.Generator().manual_seed(7)
    misc = {}
    for tag, (h, w) in {'even': (24, 32), 'odd': (27, 41)}.items():
        yuv = {'y': torch.rand(1, 1, h, w, generator=g),
               'u': torch.rand(1, 1, (h + 1) // 2, (w + 1) // 2, generator=g),
               'v': torch.rand(1, 1, (h + 1) // 2, (w + 1) // 2, generator=g)}
        x444 = rae.InputLayer()(yuv)
        assert torch.equal(x444, R.input_layer(yuv))
        o = rae.OutputLayer()(x444 * 1.3 - 0.1)
        o2 = R.output_layer(x444 * 1.3 - 0.1)
        assert all(torch.equal(o[k], o2[k]) for k in 'yuv')
        cast = cast_before_png_saving({'x': o, 'data_type': 'yuv_dic'})
        assert all(torch.equal(cast[k], R.cast_8bit(o[k])) for k in 'yuv')
        fin = R.finalize_frame(x444 * 1.3 - 0.1, h, w)
        flo = 6.0 * torch.randn(1, 2, h, w, generator=g)
        wr = ref_warp(x444, flo)
        assert torch.equal(wr, R.warp(x444, flo))
        for k in 'yuv':
            misc['%s_in_%s' % (tag, k)] = _np(yuv[k])
            misc['%s_fin_%s' % (tag, k)] = _np(fin[k])
        misc[tag + '_x444'] = _np(x444)
        misc[tag + '_flow'] = _np(flo)
        misc[tag + '_warp'] = _np(wr)
    # PdfParamParameterizer
    hs_out = 3.0 * torch.randn(1, 16, 9, 11, generator=g)
    pp = rm.PdfParamParameterizer('laplace', 8)(hs_out)
    mu, sigma = R.mu_sigma(hs_out, 8)
    assert torch.equal(pp[0]['mu'], mu) and torch.equal(pp[0]['sigma'], sigma)
    misc['hs_out'], misc['mu'], misc['sigma'] = _np(hs_out), _np(mu), _np(sigma)
    # GainMatrix
    torch.manual_seed(11)
    with contextlib.redirect_stdout(io.StringIO()):
        rg = RefGain({'N': 3, 'nb_ft': 8, 'initialize_to_one': False}).eval()
    mg = M.GainMatrix({'N': 3, 'nb_ft': 8, 'initialize_to_one': False}).eval()
    mg.load_state_dict(rg.state_dict())
    for rate in (0., 0.4, 2.0):
        for mode in ('enc', 'dec'):
            a = rg({'x': mu, 'idx_rate': rate, 'mode': mode})['output']
            assert torch.equal(a, mu * R.gain_vector(rg, rate, mode))
            assert torch.equal(a, mu * mg.gain_vector(rate, mode))
    # BallePdfEstim + ArithmeticCoder tables
    torch.manual_seed(12)
    with contextlib.redirect_stdout(io.StringIO()):
        rb = RefBalle(8, '').eval()
        ac = ArithmeticCoder({'balle_pdf_estim_z': rb, 'device': 'cpu'})
    mb = M.BallePdfEstim(8, '').eval()
    mb.load_state_dict(rb.state_dict())
    zt_ref = ac.pre_computed_z_cdf.view(8, 514)
    assert torch.equal(zt_ref, R.z_cdf_table(rb))
    assert torch.equal(zt_ref, R.z_cdf_table(mb))
    idx = torch.arange(514).float() - 256.5
    assert torch.equal(mb.cdf(idx.view(1, 1, -1, 1).repeat(1, 8, 1, 1)).view(8, 514), zt_ref)
    for k, v in state_np(rb).items():
        misc['balle_sd:' + k] = v
    misc['balle_table_u16'] = R.cdf_float_to_int(zt_ref).numpy().astype(np.uint16)
    ycdf = ac.get_y_cdf(sigma)
    assert torch.equal(ycdf, R.laplace_cdf_table(sigma))
    ref_int = R.cdf_float_to_int(ycdf).numpy().astype(np.uint16).reshape(-1, 514)
    spec_int = C.laplace_table_spec(_np(sigma).reshape(-1))
    d = np.abs(ref_int.astype(np.int32) - spec_int.astype(np.int32))
    misc['laplace_sigma'] = _np(sigma).reshape(-1)
    misc['laplace_ref_u16'] = ref_int
    misc['laplace_spec_mismatch'] = np.array([d.max(), (d != 0).mean()])
    print('laplace int-CDF: spec vs torch-fp32 reference: max |diff| = %d, mismatching entries = %.4f%%'
          % (d.max(), 100 * (d != 0).mean()))
    assert d.max() <= 1
    # GOP structures
    for n in ('1_GOP_0', 'LDP_8', '1_GOP_32', '2_GOP_16', '1_GOP_16', '3_GOP_4'):
        assert ref_gop(n) == G.generate_gop_struct(n), n
    np.savez_compressed(os.path.join(OUT, 'misc.npz'), **misc)
    print('misc ok')

    # ------------------------------------------------------------- system level
    H, W = 80, 112
    cfg = dict(seed=4321, C=32, Cy=16, Cz=16, Csc=16)
    net = models.build_standin(**cfg)
    tables = C.Tables(net)
    gop = G.generate_gop_struct('1_GOP_2')           # I0, P2, B1
    gen = torch.Generator().manual_seed(99)
    base = torch.rand(1, 1, H + 16, W + 16, generator=gen)
    base = torch.nn.functional.avg_pool2d(base, 9, stride=1, padding=4)
    base = (base - base.min()) / (base.max() - base.min())
    frames = {}
    for t in range(3):
        yv = base[:, :, t:t + H, 2 * t:2 * t + W]
        uv = torch.nn.functional.avg_pool2d(yv, 2)
        q8 = lambda a: (a * 255).round() / 255
        frames['frame_%d' % t] = {'y': q8(yv), 'u': q8(uv), 'v': q8(1 - uv)}
    sysrec = {'H': H, 'W': W, 'sd_hash': sd_hash(net)}
    for cdf_mode in ('reference', 'spec'):
        bts, rec = C.encode_gop(net, tables, frames, gop, cdf_mode=cdf_mode)
        dec = C.decode_gop(net, tables, bts, gop, H, W, cdf_mode=cdf_mode)
        for f in rec:
            assert all(torch.equal(rec[f][k], dec[f][k]) for k in 'yuv'), 'closed loop broken'
        for f in bts:
            sysrec['%s_bytes_%s' % (cdf_mode, f)] = np.frombuffer(bts[f], dtype=np.uint8)
            for k in 'yuv':
                sysrec['%s_rec_%s_%s' % (cdf_mode, f, k)] = (_np(rec[f][k]) * 255).round().astype(np.uint8)
        print('system', cdf_mode, {f: len(b) for f, b in bts.items()})
        if cdf_mode == 'reference':
            ref_bts, ref_rec = bts, rec
    for t in range(3):
        for k in 'yuv':
            sysrec['src_frame_%d_%s' % (t, k)] = (_np(frames['frame_%d' % t][k]) * 255).round().astype(np.uint8)

    # the reference's own decoder on the oracle encoder's bitstream
    for cn in (net.mode_net.mode_net, net.codec_net.codec_net):
        with contextlib.redirect_stdout(io.StringIO()):
            cn.ac = ArithmeticCoder({'balle_pdf_estim_z': cn.pdf_z, 'device': 'cpu'})
    # mirrors have no torch forward: give the reference decoder oracle-evaluated transforms
    class _Eval(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x):
            return R.forward_module(self.m, x)

    class _MC(torch.nn.Module):
        def forward(self, p):
            return {'x_warp': p['beta'] * ref_warp(p['prev'], p['v_prev'])
                    + (1 - p['beta']) * ref_warp(p['next'], p['v_next'])}

    import copy
    rnet = copy.copy(net)
    rnet._modules = dict(net._modules)
    rnet.motion_compensation = _MC()
    rnet.in_layer, rnet.out_layer = rae.InputLayer(), rae.OutputLayer()
    for wrap, attr in ((net.mode_net, 'mode_net'), (net.codec_net, 'codec_net')):
        cn = getattr(wrap, attr)
        for t in ('g_s', 'h_s', 'g_a_ref'):
            cn._modules[t] = _Eval(cn._modules[t])
        rp = rm.PdfParamParameterizer('laplace', cn.nb_ft_y)
        cn._modules['pdf_parameterizer'] = rp
        for gname in ('gain_I', 'gain_P', 'gain_B'):
            old = cn._modules[gname]
            with contextlib.redirect_stdout(io.StringIO()):
                new = RefGain({'N': 1, 'nb_ft': cn.nb_ft_y}).eval()
            new.load_state_dict(old.state_dict())
            cn._modules[gname] = new
    with contextlib.redirect_stdout(io.StringIO()):
        rdec = RefDecoder({'full_net': rnet}).eval()
    dims_y, dims_z = C.latent_dims(H, W)
    data_dim = {'x': (H, W), 'y': dims_y, 'z': dims_z, 'x_uv': ((H + 1) // 2, (W + 1) // 2)}
    decoded = {}
    with tempfile.TemporaryDirectory() as td:
        for f in sorted(gop, key=lambda f: gop[f]['coding_order']):
            path = os.path.join(td, f.split('_')[1])
            with open(path, 'wb') as fo:
                fo.write(ref_bts[f])
            t = gop[f]['type']
            prev = decoded[gop[f]['prev_ref']] if t != 0 else C.zero_yuv(H, W)
            nxt = decoded[gop[f]['next_ref']] if t == 2 else C.zero_yuv(H, W)
            with contextlib.redirect_stdout(io.StringIO()):
                out = rdec.decode({'prev_dic': prev, 'next_dic': nxt, 'frame_type': t,
                                   'bitstream_path': path, 'data_dim': data_dim, 'device': 'cpu'})
            decoded[f] = out
            for k in 'yuv':
                assert torch.equal(out[k], ref_rec[f][k]), 'reference Decoder != oracle on ' + f + k
            print('reference real_life.decode.Decoder reproduces oracle frame', f)
    np.savez_compressed(os.path.join(OUT, 'system_80x112.npz'), **sysrec)
    print('system ok ->', OUT)


if __name__ == '__main__':
    main()
