"""The ``models`` package the reference tree is missing.

``real_life/decode.py:447-453,770-795`` and ``model_management.py:97,320,351-359``
fix the attribute contract of the pickled AIVC model:

  FullNet.codec_net.codec_net   -> ConditionalNet   (CodecNet)
  FullNet.mode_net.mode_net     -> ConditionalNet   (MOFNet)
  FullNet.motion_compensation(dict) -> {'x_warp'}
  FullNet.in_layer / out_layer, FullNet.model_param['lambda_tradeoff']
  FullNet.GOP_forward(model_input) -> net_out
  ConditionalNet.{g_a, g_a_ref, h_a, h_s, g_s, pdf_y, pdf_z, pdf_parameterizer,
                  out_c_shortcut_y, nb_ft_y, nb_ft_z, gain_I, gain_P, gain_B,
                  flag_gain_p_b, ac}

The classes below honour that contract.  The transform topologies are the
documented stand-in of SURVEY.md 8(c) (Cheng-2020 style; the released weights
and the upstream topology are not in the tree), built only from the mirrored
leaf classes so that every leaf class is exercised.
"""
import torch
from torch import nn

from .layers import (CustomConvLayer, UpscalingLayer, ChengResBlock, SimplifiedAttention,
                     PdfParamParameterizer, InputLayer, OutputLayer, GainMatrix, BallePdfEstim,
                     Quantizer, ParametricPdf, EntropyCoder)

FRAME_I, FRAME_P, FRAME_B = 0, 1, 2      # func_util/GOP_structure.py:23-25


def analysis(in_c, C, out_c):
    """x -> latent at 1/16 resolution (4 stride-2 stages)."""
    return nn.Sequential(
        CustomConvLayer(5, in_c, C, non_linearity='gdn', conv_stride=2),
        ChengResBlock(C, 'plain'),
        ChengResBlock(C, 'down'),
        SimplifiedAttention(C),
        ChengResBlock(C, 'plain'),
        ChengResBlock(C, 'down'),
        ChengResBlock(C, 'plain'),
        CustomConvLayer(3, C, out_c, non_linearity='no', conv_stride=2),
        SimplifiedAttention(out_c))


def synthesis(in_c, C, out_c):
    """latent (+ shortcut) -> 16x larger output; mirror of ``analysis``."""
    return nn.Sequential(
        SimplifiedAttention(in_c),
        UpscalingLayer(3, in_c, C, non_linearity='no'),
        ChengResBlock(C, 'plain'),
        ChengResBlock(C, 'up_tconv'),
        ChengResBlock(C, 'plain'),
        SimplifiedAttention(C),
        ChengResBlock(C, 'up_tconv'),
        ChengResBlock(C, 'plain'),
        UpscalingLayer(5, C, out_c, non_linearity='no'))


def hyper_analysis(Cy, C, Cz):
    return nn.Sequential(
        CustomConvLayer(3, Cy, C, non_linearity='leaky_relu'),
        CustomConvLayer(3, C, C, non_linearity='leaky_relu', conv_stride=2),
        CustomConvLayer(3, C, Cz, non_linearity='no', conv_stride=2))


def hyper_synthesis(Cz, C, Cy):
    return nn.Sequential(
        UpscalingLayer(3, Cz, C, non_linearity='leaky_relu'),
        UpscalingLayer(3, C, C, non_linearity='leaky_relu'),
        CustomConvLayer(3, C, 2 * Cy, non_linearity='no'))



class ConditionalNet(nn.Module):
    def __init__(self, in_c, ref_c, out_c, C=128, Cy=64, Cz=64, Csc=64):
        super().__init__()
        self.in_c, self.ref_c, self.out_c = in_c, ref_c, out_c
        self.nb_ft_y, self.nb_ft_z, self.out_c_shortcut_y = Cy, Cz, Csc
        self.g_a = analysis(in_c, C, Cy)
        self.g_a_ref = analysis(ref_c, C, Csc)
        self.h_a = hyper_analysis(Cy, C, Cz)
        self.h_s = hyper_synthesis(Cz, C, Cy)
        self.g_s = synthesis(Cy + Csc, C, out_c)
        self.pdf_y = ParametricPdf('laplace')     # (inside the codec the fused quantise kernel evaluates the same rate)
        self.entropy_coder = EntropyCoder()
        self.pdf_z = BallePdfEstim(Cz, pdf_family='')
        self.pdf_parameterizer = PdfParamParameterizer('laplace', Cy)
        self.quantizer = Quantizer()
        self.flag_gain_p_b = True
        for name in ('gain_I', 'gain_P', 'gain_B'):
            setattr(self, name, GainMatrix({'N': 1, 'nb_ft': Cy, 'initialize_to_one': True}))
        self.ac = None          # attached at load time (model_management.py:351-359)


class _Wrap(nn.Module):
    def __init__(self, attr, net):
        super().__init__()
        setattr(self, attr, net)


class MotionCompensation(nn.Module):
    """x_warp = beta * warp(prev, v_prev) + (1 - beta) * warp(next, v_next)
    (contract: real_life/decode.py:524-533; warp: func_util/optical_flow.py:14-55)."""

    def forward(self, param):
        from . import ops
        return {'x_warp': ops.warp_blend(param['prev'], param['next'], param['v_prev'],
                                         param['v_next'], param['beta'])}


class FullNet(nn.Module):
    def __init__(self, C=128, Cy=64, Cz=64, Csc=64):
        super().__init__()
        self.model_param = {'lambda_tradeoff': [0.0], 'C': C, 'Cy': Cy, 'Cz': Cz, 'Csc': Csc}
        # MOFNet: in = (code, prev, next) = 9 ch, shortcut in = (prev, next) = 6, out = 6
        self.mode_net = _Wrap('mode_net', ConditionalNet(9, 6, 6, C, Cy, Cz, Csc))
        # CodecNet: in = (code, prediction) = 6 ch, shortcut in = prediction = 3, out = 3
        self.codec_net = _Wrap('codec_net', ConditionalNet(6, 3, 3, C, Cy, Cz, Csc))
        self.motion_compensation = MotionCompensation()
        self.in_layer = InputLayer()
        self.out_layer = OutputLayer()

    def GOP_forward(self, model_input):
        from .codec import gop_forward
        return gop_forward(self, model_input)


def randomize_(model, seed):
    """Make the stand-in less degenerate than default init: GDN gammas get a
    random positive off-diagonal part, gains and Balle parameters are perturbed.
    Deterministic in ``seed`` (CPU generator), independent of construction order."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith('.gamma'):
                p.add_(0.02 * torch.rand(p.shape, generator=g))
            elif name.endswith('.beta'):
                p.add_(0.1 * torch.rand(p.shape, generator=g))
        # Untrained transforms shrink their input, so every latent would round to zero and
        # the entropy-coding half of the path would be idle.  The multi-rate gains are what
        # AIVC itself uses to scale latents: give the active channels an encoder gain of
        # ~20 (decoder gain = 1/enc), leave every 4th channel "dead" (gain 1e-3, mu forced
        # to 0) so the non-zero-channel signalling of bitstream.py:241-255 is exercised, and
        # bias the log-variance half of h_s so sigma ~ e.
        for cn in (model.mode_net.mode_net, model.codec_net.codec_net):
            cy = cn.nb_ft_y
            dead = torch.arange(cy) % 4 == 3
            base = 20.0 * (0.75 + 0.5 * torch.rand(cy, generator=g))
            for k, name in enumerate(('gain_I', 'gain_P', 'gain_B')):
                gm = getattr(cn, name)
                enc = torch.where(dead, torch.full_like(base, 1e-3), base * (1.0 - 0.1 * k))
                gm.enc_gain_list[0].copy_(enc.view(cy, 1, 1))
                gm.dec_gain_list[0].copy_(torch.where(dead, torch.ones_like(enc), 1.0 / enc).view(cy, 1, 1))
            last = cn.h_s[-1].layers[1]
            last.bias[cy:2 * cy] = 2.0
            last.weight[:cy][dead] = 0.0
            last.bias[:cy][dead] = 0.0
    return model


def build_standin(seed=1234, C=128, Cy=64, Cz=64, Csc=64, hyper_boost=None):
    """Seeded stand-in FullNet (default PyTorch conv init, GDN default init, Balle
    xavier init -- SURVEY.md 8(d)) in eval mode on the CPU.
    hyper_boost=(a, s): scale the last layer of h_a by `a` and the weights of the last layer of h_s by `s`.
    Untrained hyper transforms shrink their input ~10x per layer, so without it every z rounds to zero and
    mu / sigma are spatially constant; (12, 8) gives z indices with a spread of ~1.5 and a sigma that varies
    by ~2x over the latent, i.e. a hyperprior that actually steers the range coder."""
    torch.manual_seed(seed)
    net = FullNet(C, Cy, Cz, Csc)
    randomize_(net, seed + 1)
    if hyper_boost is not None:
        a, s_ = hyper_boost
        with torch.no_grad():
            for cn in (net.mode_net.mode_net, net.codec_net.codec_net):
                last = cn.h_a[-1].layers[1]
                last.weight.mul_(a)
                last.bias.mul_(a)
                cn.h_s[-1].layers[1].weight.mul_(s_)
    return net.eval()
