"""Factories of the leaf-class golden cases (same table as oracle/gen_golden.py)."""
C16 = 16
LEAVES = {
    'conv_k5_s2_gdn': lambda L: L.CustomConvLayer(5, 9, C16, non_linearity='gdn', conv_stride=2),
    'conv_k3_s1_leaky': lambda L: L.CustomConvLayer(3, C16, C16, non_linearity='leaky_relu'),
    'conv_k3_s2_no': lambda L: L.CustomConvLayer(3, C16, 8, non_linearity='no', conv_stride=2),
    'conv_k3_s1_relu': lambda L: L.CustomConvLayer(3, C16, C16, non_linearity='relu'),
    'conv_k3_s1_igdn': lambda L: L.CustomConvLayer(3, C16, C16, non_linearity='gdn_inverse'),
    'up_k3_leaky': lambda L: L.UpscalingLayer(3, C16, C16, non_linearity='leaky_relu'),
    'up_k5_no': lambda L: L.UpscalingLayer(5, C16, 3, non_linearity='no'),
    'up_k3_igdn': lambda L: L.UpscalingLayer(3, 8, C16, non_linearity='gdn_inverse'),
    'cheng_plain': lambda L: L.ChengResBlock(C16, 'plain'),
    'cheng_down': lambda L: L.ChengResBlock(C16, 'down'),
    'cheng_up': lambda L: L.ChengResBlock(C16, 'up_tconv'),
    'resblock': lambda L: L.ResBlock(3, C16),
    'attresblock': lambda L: L.AttentionResBlock(C16),
    'attention': lambda L: L.SimplifiedAttention(C16),
    'attention_light': lambda L: L.SimplifiedAttention(C16, lightweight_resblock=True),
}


def load_leaf(name, golden_dir):
    import os
    import numpy as np
    import torch
    import aivc_b200.layers as M
    fx = np.load(os.path.join(golden_dir, 'leaf_%s.npz' % name))
    m = LEAVES[name](M).eval()
    sd = {k[3:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith('sd:')}
    m.load_state_dict(sd, strict=True)
    return m, fx
