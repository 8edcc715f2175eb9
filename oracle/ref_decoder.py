"""ORACLE (test infrastructure; needs /root/reference): the REFERENCE's own decoder -- real_life.decode.Decoder with its
ArithmeticCoder (real_life/bitstream.py; torchac replaced by oracle/torchac_shim.py) -- wrapped around a stand-in model
whose transforms are evaluated by the oracle's torch restatement (oracle/nn_ref.py; the mirrors have no torch forward).
Used by oracle/gen_golden.py (decodes the oracle encoder's streams) and oracle/check_reference_decodes.py (decodes the
CUDA encoder's streams)."""
import contextlib
import copy
import io
import os
import sys
import tempfile

import torch

REF_SRC = '/root/reference/src'


_PREFIXES = ('layers', 'models', 'real_life', 'func_util', 'model_mngt', 'torchac')


def reference_modules():
    """Import the REFERENCE's modules (torchac shimmed) and return the classes needed, whatever is registered under
    their names at the moment: aivc_b200.compat.install() aliases `layers.*` to the CUDA mirrors in sys.modules, and an
    oracle must not pick those up.  sys.modules / sys.path are restored afterwards."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    saved = {k: v for k, v in sys.modules.items() if k.split('.')[0] in _PREFIXES}
    for k in saved:
        del sys.modules[k]
    old_path = list(sys.path)
    sys.path[:0] = [REF_SRC, root]
    try:
        from oracle import torchac_shim
        sys.modules['torchac'] = torchac_shim
        with contextlib.redirect_stdout(io.StringIO()):
            from layers.misc import misc_layers as rm
            from layers.ae import ae_layers as rae
            from layers.multi_rate.gain_matrix import GainMatrix as RefGain
            from func_util.optical_flow import warp as ref_warp
            from real_life.bitstream import ArithmeticCoder
            from real_life.decode import Decoder as RefDecoder
        assert rm.__file__.startswith(REF_SRC), rm.__file__
        return dict(rm=rm, rae=rae, RefGain=RefGain, ref_warp=ref_warp, ArithmeticCoder=ArithmeticCoder,
                    RefDecoder=RefDecoder)
    finally:
        for k in [k for k in sys.modules if k.split('.')[0] in _PREFIXES]:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.path[:] = old_path


def build_reference_decoder(net):
    """real_life.decode.Decoder({'full_net': ...}) for stand-in `net` (left untouched: a shallow copy is rewired)."""
    from oracle import nn_ref as R
    ns = reference_modules()
    ref_warp = ns['ref_warp']

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

    rnet = copy.copy(net)
    rnet._modules = dict(net._modules)
    rnet.motion_compensation = _MC()
    rnet.in_layer, rnet.out_layer = ns['rae'].InputLayer(), ns['rae'].OutputLayer()
    for attr in ('mode_net', 'codec_net'):
        wrap = copy.copy(getattr(net, attr))
        wrap._modules = dict(wrap._modules)
        cn = copy.copy(getattr(wrap, attr))
        cn._modules = dict(cn._modules)
        with contextlib.redirect_stdout(io.StringIO()):
            cn.ac = ns['ArithmeticCoder']({'balle_pdf_estim_z': cn.pdf_z, 'device': 'cpu'})
        for t in ('g_s', 'h_s', 'g_a_ref'):
            cn._modules[t] = _Eval(cn._modules[t])
        cn._modules['pdf_parameterizer'] = ns['rm'].PdfParamParameterizer('laplace', cn.nb_ft_y)
        for gname in ('gain_I', 'gain_P', 'gain_B'):
            old = cn._modules[gname]
            with contextlib.redirect_stdout(io.StringIO()):
                new = ns['RefGain']({'N': 1, 'nb_ft': cn.nb_ft_y}).eval()
            new.load_state_dict(old.state_dict())
            cn._modules[gname] = new
        wrap._modules[attr] = cn
        rnet._modules[attr] = wrap
    with contextlib.redirect_stdout(io.StringIO()):
        return ns['RefDecoder']({'full_net': rnet}).eval()


def reference_decode_gop(rdec, frame_bytes, gop, h, w):
    """decode_one_GOP's loop (real_life/decode.py:244-301) on in-memory frame bitstreams -> {'frame_i': YUV420 dict}."""
    from oracle import codec_ref as C
    dims_y, dims_z = C.latent_dims(h, w)
    data_dim = {'x': (h, w), 'y': dims_y, 'z': dims_z, 'x_uv': ((h + 1) // 2, (w + 1) // 2)}
    decoded = {}
    with tempfile.TemporaryDirectory() as td, torch.no_grad():
        for f in sorted(gop, key=lambda f: gop[f]['coding_order']):
            path = os.path.join(td, f.split('_')[1])
            with open(path, 'wb') as fo:
                fo.write(frame_bytes[f])
            t = gop[f]['type']
            prev = decoded[gop[f]['prev_ref']] if t != 0 else C.zero_yuv(h, w)
            nxt = decoded[gop[f]['next_ref']] if t == 2 else C.zero_yuv(h, w)
            with contextlib.redirect_stdout(io.StringIO()):
                decoded[f] = rdec.decode({'prev_dic': prev, 'next_dic': nxt, 'frame_type': t, 'bitstream_path': path,
                                          'data_dim': data_dim, 'device': 'cpu'})
    return decoded
