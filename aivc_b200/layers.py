"""Drop-in mirrors of the AIVC layer classes, executed by the B200 engine.

Every class here has the same name, constructor signature, sub-module layout
and therefore the same ``state_dict`` keys as the reference class it replaces,
so weights move between the two with ``load_state_dict`` and a pickled AIVC
model can be re-pointed at these classes (see ``aivc_b200.compat``).  The
classes only *hold* parameters; ``forward`` lowers the module (once) to a fused
kernel plan (``aivc_b200.plan``) and launches hand-written sm_100a kernels
through the C-ABI library.  There is no PyTorch / CPU fallback: without the
CUDA library ``forward`` raises.

Reference classes mirrored (file:line under /root/reference/src):
  CustomConvLayer     layers/misc/custom_conv_layers.py:129-180
  UpscalingLayer      layers/misc/custom_conv_layers.py:183-253
  ChengResBlock       layers/misc/custom_conv_layers.py:21-109
  ResBlock            layers/misc/custom_conv_layers.py:112-126
  GDN                 layers/misc/misc_layers.py:63-154
  Quantizer           layers/misc/misc_layers.py:157-169
  PdfParamParameterizer layers/misc/misc_layers.py:172-269
  AttentionResBlock / SimplifiedAttention  layers/misc/attention.py:22-97
  InputLayer / OutputLayer  layers/ae/ae_layers.py:17-56
  GainMatrix          layers/multi_rate/gain_matrix.py:27-194
  BallePdfEstim       layers/entropy_coding/pdf_estimator.py:73-245
"""
import math

import torch
from torch import nn

LOG_VAR_MIN = -18.4207   # func_util/math_func.py:31
LOG_VAR_MAX = 10.0       # func_util/math_func.py:30

_ACTS = ('gdn', 'gdn_inverse', 'leaky_relu', 'relu', 'no')


def _xavier(shape):
    """N(0, sqrt(2 / numel)) init, func_util/math_func.py:34-50."""
    shape = torch.Size(shape)
    return torch.randn(shape) * math.sqrt(2.0 / shape.numel())


class _Engine(nn.Module):
    """Base: lowers ``self`` to a kernel plan on first call and runs it."""

    def forward(self, x):
        from . import plan
        return plan.run_module(self, x)


class GDN(_Engine):
    """y_i = x_i * (beta_i + sum_j gamma_ij x_j^2)^(-1/2)  (or ^(+1/2) if inverse)."""

    def __init__(self, ch, inverse=False, beta_min=1e-6, gamma_init=.1,
                 reparam_offset=2 ** -18):
        super().__init__()
        self.inverse = bool(inverse)
        # plain attributes (not buffers), as in the reference (misc_layers.py:78-111)
        self.reparam_offset = torch.tensor([reparam_offset], dtype=torch.float32)
        self.pedestal = self.reparam_offset ** 2
        self.beta_bound = (beta_min + self.reparam_offset ** 2) ** .5
        self.gamma_bound = self.reparam_offset
        self.beta = nn.Parameter(torch.sqrt(torch.ones(ch) + self.pedestal))
        self.gamma = nn.Parameter(torch.sqrt(gamma_init * torch.eye(ch) + self.pedestal))

    def effective(self):
        """(beta[C], gamma[C_out, C_in]) after the lower-bound reparametrisation
        (misc_layers.py:131-139), fp32, computed once at plan-build time."""
        dev = self.beta.device
        ped = self.pedestal.to(dev)
        beta = torch.maximum(self.beta.detach(), self.beta_bound.to(dev)) ** 2 - ped
        gamma = torch.maximum(self.gamma.detach(), self.gamma_bound.to(dev)) ** 2 - ped
        return beta.float().contiguous(), gamma.float().contiguous()


def _add_act(seq, non_linearity, out_ft):
    if non_linearity == 'gdn':
        seq.add_module('non_linearity', GDN(out_ft, inverse=False))
    elif non_linearity == 'gdn_inverse':
        seq.add_module('non_linearity', GDN(out_ft, inverse=True))
    elif non_linearity == 'leaky_relu':
        seq.add_module('non_linearity', nn.LeakyReLU())
    elif non_linearity == 'relu':
        seq.add_module('non_linearity', nn.ReLU())


class CustomConvLayer(_Engine):
    def __init__(self, k_size=5, in_ft=64, out_ft=64, flag_bias=True,
                 non_linearity='leaky_relu', conv_stride=1, padding_mode='replicate'):
        super().__init__()
        if padding_mode != 'replicate':
            raise ValueError('only replicate padding exists in AIVC')
        self.non_linearity = non_linearity
        self.layers = nn.Sequential(
            nn.ReplicationPad2d(k_size // 2),
            nn.Conv2d(in_ft, out_ft, k_size, stride=conv_stride, bias=flag_bias))
        _add_act(self.layers, non_linearity, out_ft)


class UpscalingLayer(_Engine):
    def __init__(self, k_size=5, in_ft=64, out_ft=64, flag_bias=True,
                 non_linearity='leaky_relu', mode='transposed', flag_first_layer=False):
        super().__init__()
        if mode == 'transposed_no_bias':
            flag_bias = False
        self.non_linearity = non_linearity
        pad = int(((1 + k_size) / 2) - 1)
        self.layers = nn.Sequential(
            nn.ConvTranspose2d(in_ft, out_ft, k_size, stride=2, padding=pad,
                               output_padding=1, bias=flag_bias))
        _add_act(self.layers, non_linearity, out_ft)


class ChengResBlock(_Engine):
    def __init__(self, nb_ft, mode='plain'):
        super().__init__()
        self.mode = mode
        if mode == 'plain':
            self.layers = nn.Sequential(
                CustomConvLayer(3, nb_ft, nb_ft, non_linearity='leaky_relu'),
                CustomConvLayer(3, nb_ft, nb_ft, non_linearity='leaky_relu'))
        elif mode == 'down':
            self.layers = nn.Sequential(
                CustomConvLayer(3, nb_ft, nb_ft, non_linearity='leaky_relu', conv_stride=2),
                CustomConvLayer(3, nb_ft, nb_ft, non_linearity='gdn'))
            self.aux_layer = nn.Conv2d(nb_ft, nb_ft, 1, stride=2)
        elif mode == 'up_tconv':
            self.layers = nn.Sequential(
                UpscalingLayer(3, nb_ft, nb_ft, non_linearity='leaky_relu'),
                CustomConvLayer(3, nb_ft, nb_ft, non_linearity='gdn_inverse'))
            self.aux_layer = UpscalingLayer(3, nb_ft, nb_ft, non_linearity='no')
        else:
            raise ValueError(mode)


class ResBlock(_Engine):
    def __init__(self, k_size, nb_ft):
        super().__init__()
        p = k_size // 2
        self.layers = nn.Sequential(
            nn.ReplicationPad2d(p), nn.Conv2d(nb_ft, nb_ft, k_size), nn.ReLU(),
            nn.ReplicationPad2d(p), nn.Conv2d(nb_ft, nb_ft, k_size))


class AttentionResBlock(_Engine):
    def __init__(self, nb_ft):
        super().__init__()
        half = nb_ft // 2
        self.layers = nn.Sequential(
            nn.Conv2d(nb_ft, half, 1), nn.LeakyReLU(),
            nn.ReplicationPad2d(1), nn.Conv2d(half, half, 3), nn.LeakyReLU(),
            nn.Conv2d(half, nb_ft, 1))


class SimplifiedAttention(_Engine):
    def __init__(self, nb_ft, k_size=3, lightweight_resblock=False):
        super().__init__()
        self.nb_ft, self.k_size = nb_ft, k_size
        mk = (lambda: AttentionResBlock(nb_ft)) if lightweight_resblock \
            else (lambda: ResBlock(k_size, nb_ft))
        self.trunk = nn.Sequential(mk(), mk(), mk())
        self.attention = nn.Sequential(mk(), mk(), mk(),
                                       nn.Conv2d(nb_ft, nb_ft, 1), nn.Sigmoid())


class Quantizer(nn.Module):
    """Inference-time quantiser: round-half-to-even (misc_layers.py:167)."""

    def forward(self, x, fine_tune=False):
        if self.training or fine_tune:
            return x + (torch.rand_like(x) - 0.5)
        return torch.round(x)


class PdfParamParameterizer(nn.Module):
    """Splits the hyper-decoder output into the parameters of a K-component mixture (misc_layers.py:180-269):
    channels [mu_0..mu_K-1 | log var_0.. | (log gamma_0.. if 'gamma') | weight logits_1..K-1], sigma =
    exp(0.5 clamp(log var, LOG_VAR_MIN, LOG_VAR_MAX)), weights = softmax over (1, logits).  AIVC ships K = 1."""

    def __init__(self, ec_mode, nb_ft):
        super().__init__()
        self.ec_mode, self.nb_ft = ec_mode, nb_ft

    def forward(self, x):
        from . import ops
        toks = self.ec_mode.split('_')
        K = 2 if 'two' in toks else (3 if 'three' in toks else 1)
        C = self.nb_ft
        if K == 1 and 'gamma' not in toks:
            mu, sigma = ops.mu_sigma(x, C)
            return [{'mu': mu, 'sigma': sigma, 'gamma': torch.ones_like(mu), 'weight': torch.ones_like(mu)}]
        sl = lambda i: x[:, i * C:(i + 1) * C]
        comp = [ops.mu_sigma(torch.cat((sl(k), sl(K + k)), 1), C) for k in range(K)]
        pos = 2 * K
        gammas = []
        for k in range(K):
            if 'gamma' in toks:
                gammas.append(ops.mu_sigma(torch.cat((sl(pos), sl(pos)), 1), C)[1])
                pos += 1
            else:
                gammas.append(torch.ones_like(comp[k][0]))
        w = torch.stack([torch.ones_like(comp[0][0])] + [sl(pos + k - 1).float() for k in range(1, K)], 1)
        w = torch.softmax(w, dim=1)
        return [{'mu': comp[k][0], 'sigma': comp[k][1], 'gamma': gammas[k], 'weight': w[:, k]} for k in range(K)]


class ParametricPdf(nn.Module):
    """P(y) = sum over components of cdf(y + 1/2) - cdf(y - 1/2), Laplace(mu, sigma / sqrt 2) or Normal(mu, sigma)
    (pdf_estimator.py:17-70; like the reference, component weights are NOT applied here)."""

    def __init__(self, pdf_family):
        super().__init__()
        self.pdf_family = pdf_family

    def forward(self, y_tilde, all_pdf_param, zero_mu=False):
        from . import ops
        toks = self.pdf_family.split('_')
        family = 'normal' if 'normal' in toks else 'laplace'
        out = None
        for prm in all_pdf_param:
            mu = None if ('mu' in toks or zero_mu) else prm.get('mu')
            out = ops.pdf_prob(y_tilde, mu, prm.get('sigma'), family, out)
        return out.view_as(y_tilde)


class EntropyCoder(nn.Module):
    """rate = -log2 clamp(p, 2^-16, 1)   (entropy_coder.py:18-30)"""

    def forward(self, prob_x, x=None):
        return -torch.log2(torch.clamp(prob_x, 2.0 ** -16, 1.0))


class View(nn.Module):
    """misc_layers.py:30-36 (appears in pickled models; pure reshape)"""

    def __init__(self, shape):
        super().__init__()
        self.shape = shape

    def forward(self, x):
        return x.view(*self.shape)


class LowerBound(torch.autograd.Function):
    """max(inputs, bound) with the pass-through gradient of misc_layers.py:39-60 (GDN's reparametrisation; the
    inference path folds it into the packed GDN parameters, plan._gdn_params)."""

    @staticmethod
    def forward(ctx, inputs, bound):
        b = torch.ones_like(inputs) * bound
        ctx.save_for_backward(inputs, b)
        return torch.max(inputs, b)

    @staticmethod
    def backward(ctx, grad_output):
        inputs, b = ctx.saved_tensors
        return ((inputs >= b) | (grad_output < 0)).type(grad_output.dtype) * grad_output, None


class InputLayer(nn.Module):
    """YUV420 dict -> 3-channel 4:4:4 tensor (nearest x2 of U,V) ae_layers.py:27-35."""

    def forward(self, x):
        from . import ops
        return ops.yuv420_to_444(x['y'], x['u'], x['v'])


class OutputLayer(nn.Module):
    """4:4:4 tensor -> YUV420 dict (bilinear x0.5 of ch 1,2) ae_layers.py:42-56."""

    def __init__(self, k_size=5):
        super().__init__()

    def forward(self, x):
        from . import ops
        y, u, v = ops.yuv444_to_420(x)
        return {'y': y, 'u': u, 'v': v}


class GainMatrix(nn.Module):
    """Per-channel |gain| vectors, geometric interpolation on a float rate index
    (gain_matrix.py:92-194).  Dict in, dict out, like the reference."""

    def __init__(self, param):
        super().__init__()
        n, nb_ft = param.get('N'), param.get('nb_ft')
        to_one = param.get('initialize_to_one', True)
        dim = (1, 1, 1) if param.get('scalar_gain', False) else (nb_ft, 1, 1)
        self.enc_gain_list = nn.ParameterList()
        self.dec_gain_list = nn.ParameterList()
        for _ in range(n):
            self.enc_gain_list.append(nn.Parameter(torch.ones(dim) if to_one else _xavier(dim)))
            self.dec_gain_list.append(nn.Parameter(torch.ones(dim) if to_one else _xavier(dim)))

    def gain_vector(self, idx_rate, mode):
        lst = {'enc': self.enc_gain_list, 'dec': self.dec_gain_list}[mode]
        if self.training:
            return lst[int(idx_rate)].detach().abs()
        lo = int(math.floor(idx_rate))
        hi = lo + 1
        lam = 1 - (idx_rate - lo)
        if hi == len(lst):
            hi = lo
        return (lst[lo].detach().abs() ** lam) * (lst[hi].detach().abs() ** (1 - lam))

    def forward(self, param):
        from . import ops
        g = self.gain_vector(param.get('idx_rate', 0.), param.get('mode'))
        return {'output': ops.channel_scale(param.get('x'), g)}


class BallePdfEstim(nn.Module):
    """Factorised-prior CDF network (K=4, r=3).  Only evaluated once per model
    load to build the 514-entry z table, so it stays a host-side torch module."""

    def __init__(self, nb_channel, pdf_family='', verbose=True):
        super().__init__()
        self.nb_channel, self.pdf_family = nb_channel, pdf_family
        self.K, self.r = 4, 3
        c, r = nb_channel, self.r
        corr = math.sqrt(float(c))
        self.matrix_h, self.bias_b, self.bias_a = (nn.ParameterList() for _ in range(3))
        for i in range(self.K):
            if i == 0:
                self.matrix_h.append(nn.Parameter(_xavier((c, 1, r)) * corr))
                self.bias_a.append(nn.Parameter(_xavier((c, r)) * corr))
                self.bias_b.append(nn.Parameter(_xavier((c, r)) * corr))
            elif i == self.K - 1:
                self.matrix_h.append(nn.Parameter(_xavier((c, r, 1)) * corr))
                self.bias_b.append(nn.Parameter(_xavier((c, 1)) * corr))
            else:
                self.matrix_h.append(nn.Parameter(_xavier((c, r, r)) * corr))
                self.bias_a.append(nn.Parameter(_xavier((c, r)) * corr))
                self.bias_b.append(nn.Parameter(_xavier((c, r)) * corr))

    def cdf(self, x):
        """x: [B, C, E, 1] -> CDF values [B, C, E, 1] (pdf_estimator.py:204-245)."""
        t = x
        for i in range(self.K):
            t = torch.einsum('bced,cdr->bcer', t, nn.functional.softplus(self.matrix_h[i]))
            t = t + self.bias_b[i].unsqueeze(1)
            if i != self.K - 1:
                t = t + torch.tanh(self.bias_a[i].unsqueeze(1)) * torch.tanh(t)
        return torch.sigmoid(t)

    def forward(self, x_tilde, pdf_param=None):
        b, c, h, w = x_tilde.shape
        x = x_tilde.reshape(b, c, h * w, 1)
        return (self.cdf(x + 0.5) - self.cdf(x - 0.5)).view(b, c, h, w)
