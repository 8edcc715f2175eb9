"""ORACLE (test infrastructure only -- never imported by the product path).

CPU fp32 restatement of the AIVC layer arithmetic, written functionally on top of
``torch.nn.functional`` and driven by *class names*, so it interprets both the
real reference modules (imported from /root/reference/src when generating the
golden fixtures) and the parameter-holding mirrors in ``aivc_b200.layers``.

Pinned by ``oracle/gen_golden.py``: every function below is checked bit-for-bit
against the reference class it restates, executed by the same torch build, and the
resulting input/output vectors are committed under ``tests/golden``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.
"""
import torch
import torch.nn.functional as F

LOG_VAR_MIN = -18.4207   # func_util/math_func.py:31
LOG_VAR_MAX = 10.0       # func_util/math_func.py:30


# ----------------------------------------------------------------------------- leaves
def replicate_pad(x, p):
    """ReplicationPad2d(p) -- custom_conv_layers.py:146-148."""
    return F.pad(x, (p, p, p, p), mode='replicate') if p > 0 else x


def gdn(x, beta_raw, gamma_raw, beta_bound, gamma_bound, pedestal, inverse):
    """GDN.forward -- misc_layers.py:113-154.
    beta = max(beta, beta_bound)^2 - pedestal; gamma likewise; norm = sqrt(conv1x1(x^2, gamma) + beta);
    y = x * norm if inverse else x / norm."""
    beta = torch.maximum(beta_raw, beta_bound.expand_as(beta_raw)) ** 2 - pedestal
    gamma = torch.maximum(gamma_raw, gamma_bound.expand_as(gamma_raw)) ** 2 - pedestal
    c = x.shape[1]
    norm = torch.sqrt(F.conv2d(x ** 2, gamma.view(c, c, 1, 1), bias=beta))
    return x * norm if inverse else x / norm


def _gdn_module(m, x):
    return gdn(x, m.beta.detach(), m.gamma.detach(), m.beta_bound, m.gamma_bound, m.pedestal,
               m.inverse)


def _conv(m, x):
    return F.conv2d(x, m.weight.detach(), None if m.bias is None else m.bias.detach(),
                    stride=m.stride, padding=m.padding)


def _tconv(m, x):
    return F.conv_transpose2d(x, m.weight.detach(), None if m.bias is None else m.bias.detach(),
                              stride=m.stride, padding=m.padding, output_padding=m.output_padding)


def forward_module(m, x):
    """Evaluate module ``m`` (reference class or mirror) on ``x`` [1,C,H,W] fp32 CPU."""
    name = type(m).__name__
    if name == 'Sequential':
        for child in m:
            x = forward_module(child, x)
        return x
    if name == 'ReplicationPad2d':
        return replicate_pad(x, m.padding[0])
    if name == 'Conv2d':
        return _conv(m, x)
    if name == 'ConvTranspose2d':
        return _tconv(m, x)
    if name == 'LeakyReLU':
        return F.leaky_relu(x, m.negative_slope)
    if name == 'ReLU':
        return F.relu(x)
    if name == 'Sigmoid':
        return torch.sigmoid(x)
    if name == 'GDN':
        return _gdn_module(m, x)
    if name in ('CustomConvLayer', 'UpscalingLayer'):
        # custom_conv_layers.py:179 / :252 -- just the inner Sequential
        return forward_module(m.layers, x)
    if name == 'ChengResBlock':
        # custom_conv_layers.py:105-109
        if m.mode == 'plain':
            return x + forward_module(m.layers, x)
        return forward_module(m.aux_layer, x) + forward_module(m.layers, x)
    if name == 'ResBlock':
        # custom_conv_layers.py:125-126
        return F.relu(x + forward_module(m.layers, x))
    if name == 'AttentionResBlock':
        # attention.py:41-42
        return F.leaky_relu(x + forward_module(m.layers, x))
    if name == 'SimplifiedAttention':
        # attention.py:90-97
        return forward_module(m.trunk, x) * forward_module(m.attention, x) + x
    raise NotImplementedError('oracle: no restatement for ' + name)


# ----------------------------------------------------------------------------- pixel ends
def input_layer(yuv):
    """InputLayer.forward -- ae_layers.py:27-35: Y, nearest x2 of (U,V) cropped to Y."""
    y, u, v = yuv['y'], yuv['u'], yuv['v']
    uv = F.interpolate(torch.cat((u, v), dim=1), scale_factor=2, mode='nearest')
    return torch.cat((y, uv[:, :, :y.shape[2], :y.shape[3]]), dim=1)


def output_layer(x):
    """OutputLayer.forward -- ae_layers.py:42-56."""
    uv = F.interpolate(x[:, 1:, :, :], scale_factor=0.5, mode='bilinear', align_corners=False,
                       recompute_scale_factor=False)
    return {'y': x[:, 0:1], 'u': uv[:, 0:1], 'v': uv[:, 1:2]}


def cast_8bit(x):
    """cast_before_png_saving -- func_util/img_processing.py:68-73."""
    return ((255. * torch.clamp(x, 0., 1.)).round() / 255.).float()


def finalize_frame(x444, h, w):
    """decode.py:553-577: out_layer, replicate-pad U/V to ceil(h/2) x ceil(w/2), crop, 8-bit cast."""
    o = output_layer(x444)
    h_uv, w_uv = (h + 1) // 2, (w + 1) // 2
    pr = abs(h_uv - o['u'].shape[2])
    pc = abs(w_uv - o['u'].shape[3])
    pad = lambda t: F.pad(t, (0, pc, 0, pr), mode='replicate') if (pr or pc) else t
    return {'y': cast_8bit(o['y'][:, :, :h, :w]),
            'u': cast_8bit(pad(o['u'])[:, :, :h_uv, :w_uv]),
            'v': cast_8bit(pad(o['v'])[:, :, :h_uv, :w_uv])}


def warp(x, flo):
    """func_util/optical_flow.py:14-55: bilinear grid_sample, border padding,
    align_corners=True, times the (always-one) validity mask."""
    _, _, h, w = x.shape
    xx = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w).expand(1, 1, h, w)
    yy = torch.arange(h, dtype=torch.float32).view(1, 1, h, 1).expand(1, 1, h, w)
    vx = 2.0 * (xx + flo[:, 0:1]) / max(w - 1, 1) - 1.0
    vy = 2.0 * (yy + flo[:, 1:2]) / max(h - 1, 1) - 1.0
    grid = torch.cat((vx, vy), dim=1).permute(0, 2, 3, 1)
    out = F.grid_sample(x, grid, mode='bilinear', padding_mode='border', align_corners=True)
    mask = F.grid_sample(torch.ones_like(x), grid, mode='bilinear', padding_mode='border',
                         align_corners=True)
    mask = torch.where(mask < 0.9999, torch.zeros_like(mask), torch.ones_like(mask))
    return out * mask


def motion_compensation(prev, nxt, v_prev, v_next, beta):
    """Contract of the missing models.motion_compensation (decode.py:524-533):
    x_warp = beta * warp(prev, v_prev) + (1 - beta) * warp(next, v_next)."""
    return beta * warp(prev, v_prev) + (1 - beta) * warp(nxt, v_next)


def mofnet_post(raw, h, w, frame_is_p):
    """MOFNetDecoder.decode post-processing -- decode.py:729-739."""
    o = raw[:, :, :h, :w]
    alpha = torch.clamp(o[:, 0:1] + 0.5, 0., 1.).repeat(1, 3, 1, 1)
    beta = torch.clamp(o[:, 1:2] + 0.5, 0., 1.).repeat(1, 3, 1, 1)
    v_prev, v_next = o[:, 2:4].clone(), o[:, 4:6].clone()
    if frame_is_p:
        beta = torch.ones_like(beta)
        v_next = torch.zeros_like(v_next)
    return alpha, beta, v_prev, v_next


# ----------------------------------------------------------------------------- entropy model
def mu_sigma(x, nb_ft):
    """PdfParamParameterizer.forward, K=1 -- misc_layers.py:180-269."""
    mu = x[:, :nb_ft]
    sigma = torch.exp(0.5 * torch.clamp(x[:, nb_ft:2 * nb_ft], min=LOG_VAR_MIN, max=LOG_VAR_MAX))
    return mu, sigma


def gain_vector(gm, idx_rate, mode):
    """GainMatrix.interpolate_gain_vector (eval) -- gain_matrix.py:158-194."""
    import math
    lst = gm.enc_gain_list if mode == 'enc' else gm.dec_gain_list
    lo = int(math.floor(idx_rate))
    hi = lo + 1
    lam = 1 - (idx_rate - lo)
    if hi == len(lst):
        hi = lo
    return (lst[lo].detach().abs() ** lam) * (lst[hi].detach().abs() ** (1 - lam))


def balle_cdf(pdf_z, x):
    """BallePdfEstim.cdf -- pdf_estimator.py:204-245.  x: [B,C,E,1]."""
    t = x
    k = len(pdf_z.matrix_h)
    for i in range(k):
        t = torch.einsum('bced,cdr->bcer', t, F.softplus(pdf_z.matrix_h[i].detach()))
        t = t + pdf_z.bias_b[i].detach().repeat(1, t.shape[2]).view(t.shape[1:])
        if i != k - 1:
            t = t + torch.tanh(pdf_z.bias_a[i].detach().repeat(1, t.shape[2]).view(t.shape[1:])) \
                * torch.tanh(t)
    return torch.sigmoid(t)


def z_cdf_table(pdf_z, ac_max_val=256):
    """ArithmeticCoder._precompute_z_cdf -- bitstream.py:82-125 -> float [C, 514]."""
    lp = 2 * ac_max_val + 2
    idx = torch.arange(lp).float() - ac_max_val - 0.5
    idx = idx.view(1, 1, -1, 1).repeat(1, pdf_z.nb_channel, 1, 1)
    return balle_cdf(pdf_z, idx).squeeze(-1).squeeze(0)


def laplace_cdf_table(sigma, ac_max_val=256):
    """ArithmeticCoder.get_y_cdf -- bitstream.py:127-154 -> float [..., 514].
    Laplace(0, sigma/sqrt(2)).cdf(t) = 0.5 - 0.5*sign(t)*expm1(-|t|/b) (torch.distributions)."""
    lp = 2 * ac_max_val + 2
    idx = torch.arange(lp).float() - ac_max_val - 0.5
    b = sigma.unsqueeze(-1) / torch.sqrt(torch.tensor([2.0]))
    t = idx.view(*([1] * sigma.dim()), -1)
    return 0.5 - 0.5 * torch.sign(t) * torch.expm1(-t.abs() / b)


def cdf_float_to_int(cdf_float):
    """torchac (PyPI, un-pinned; README.md:140) ``_convert_to_int_and_normalize`` with
    needs_normalization=True, as called at bitstream.py:281,454: round(cdf * (2^16 - (Lp-1)))
    -> int16 (wrapping) + arange(Lp), read back as uint16 by the C++ backend.
    parity unpinned (torchac source is not in the tree); returns int64 values in [0, 65535]."""
    lp = cdf_float.shape[-1]
    scaled = (cdf_float * float(65536 - (lp - 1))).round().to(torch.int64)
    return (scaled + torch.arange(lp, dtype=torch.int64)) & 0xFFFF


def laplace_rate_bits(q, sigma):
    """Encoder-side rate estimate (pdf_estimator.py:27-70 zero_mu + entropy_coder.py:25-30)."""
    b = sigma / torch.sqrt(torch.tensor([2.0]))
    cdf = lambda t: 0.5 - 0.5 * torch.sign(t) * torch.expm1(-t.abs() / b)
    p = torch.clamp(cdf(q + 0.5) - cdf(q - 0.5), 2.0 ** -16, 1.0)
    return -torch.log2(p)
