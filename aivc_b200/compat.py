"""Glue that lets the reference's own scripts and pickles run on this engine.

  install()        -- (a) registers the mirrors under the reference's module paths
                      (``layers.misc.custom_conv_layers`` ...; a whole-module pickle stores
                      ``module path + class name``, model_management.py:347) and the missing
                      ``models`` package; (b) provides ``torchac`` (bitstream.py:10) backed by the
                      C++ range coder; (c) restores the two torch APIs the reference needs and
                      torch >= 2 dropped.  Subprocess-safe when called from a
                      ``sitecustomize`` on PYTHONPATH.
  convert(model)   -- swaps reference layer instances of an already loaded model for mirrors
                      (weights shared, not copied), so every ``forward`` runs the CUDA path.
"""
import os
import sys
import types

import numpy as np
import torch

from . import layers as L, models as M

_MODULE_MAP = {
    'layers.misc.custom_conv_layers': ('ChengResBlock', 'ResBlock', 'CustomConvLayer', 'UpscalingLayer'),
    'layers.misc.misc_layers': ('GDN', 'Quantizer', 'PdfParamParameterizer', 'View', 'LowerBound'),
    'layers.misc.attention': ('AttentionResBlock', 'SimplifiedAttention'),
    'layers.ae.ae_layers': ('InputLayer', 'OutputLayer'),
    'layers.multi_rate.gain_matrix': ('GainMatrix',),
    'layers.entropy_coding.pdf_estimator': ('BallePdfEstim', 'ParametricPdf'),
    'layers.entropy_coding.entropy_coder': ('EntropyCoder',),
}


class _TorchacShim(types.ModuleType):
    """encode_float_cdf / decode_float_cdf with torchac's signatures, on the product coder."""

    @staticmethod
    def _to_int(cdf_float, needs_normalization):
        lp = cdf_float.shape[-1]
        if needs_normalization:
            t = (cdf_float * float(65536 - (lp - 1))).round().to(torch.int64) + torch.arange(lp)
        else:
            t = (cdf_float * 65536.0).round().to(torch.int64)
        return (t & 0xFFFF).numpy().astype(np.uint32).reshape(-1, lp)

    def encode_float_cdf(self, cdf_float, sym, needs_normalization=True, check_input_bounds=False):
        from .entropy import _encode_bounds
        lp = cdf_float.shape[-1]
        if check_input_bounds and (cdf_float.min() < 0 or cdf_float.max() > 1 or sym.max() >= lp - 1
                                   or sym.min() < 0):
            raise ValueError('torchac: input out of bounds')
        tab = self._to_int(cdf_float.cpu(), needs_normalization)
        s = sym.reshape(-1).cpu().numpy().astype(np.int64)
        ar = np.arange(s.size)
        lo = tab[ar, s]
        hi = np.where(s == lp - 2, 0x10000, tab[ar, np.minimum(s + 1, lp - 1)])
        if (hi > 0xFFFF).any():
            raise ValueError('top symbol not representable in packed bounds')
        return _encode_bounds(np.ascontiguousarray(lo | (hi << 16), dtype=np.uint32))

    def decode_float_cdf(self, cdf_float, byte_stream, needs_normalization=True):
        import ctypes as C
        from . import _lib
        lp = cdf_float.shape[-1]
        if lp != 514:
            raise ValueError('only AIVC tables (Lp = 514) are supported')
        tab = np.ascontiguousarray(self._to_int(cdf_float.cpu(), needs_normalization).astype(np.uint16))
        n = tab.shape[0]
        out = np.empty(n, dtype=np.int16)
        buf = np.frombuffer(byte_stream, dtype=np.uint8) if len(byte_stream) else np.zeros(1, np.uint8)
        # per-symbol tables: decode one "channel" of one symbol at a time is the generic form;
        # the table decoder takes [c][514] with hw symbols per row, so feed rows as channels
        _lib.check(_lib.lib().aivc_rc_decode_table(tab.ctypes.data, buf.ctypes.data, len(byte_stream),
                                                   n, 1, out.ctypes.data))
        return torch.from_numpy(out.astype(np.int16) + 256).view(cdf_float.shape[:-1])


def install():
    for path, names in _MODULE_MAP.items():
        mod = sys.modules.get(path) or types.ModuleType(path)
        for n in names:
            setattr(mod, n, getattr(L, n))
        sys.modules[path] = mod
        parts = path.split('.')
        for i in range(1, len(parts)):
            sys.modules.setdefault('.'.join(parts[:i]), types.ModuleType('.'.join(parts[:i])))
    sys.modules.setdefault('models', M)
    sys.modules.setdefault('torchac', _TorchacShim('torchac'))
    if not hasattr(torch, 'set_deterministic'):
        torch.set_deterministic = lambda flag=True: torch.use_deterministic_algorithms(flag, warn_only=True)
    if not getattr(torch.load, '_aivc_patched', False):
        _orig = torch.load

        def load(*a, **k):
            # AIVC checkpoints are whole-module pickles (model_management.py:347 calls torch.load(path, map_location=..)).
            # Only THAT call site -- and callers that set AIVC_B200_TRUST_PICKLES=1 -- get the unsafe default back;
            # everybody else keeps torch's weights_only=True.
            caller = sys._getframe(1).f_globals.get('__name__', '')
            if caller.endswith('model_management') or os.environ.get('AIVC_B200_TRUST_PICKLES') == '1':
                k.setdefault('weights_only', False)
            return _orig(*a, **k)
        load._aivc_patched = True
        torch.load = load


def _mirror_of(m):
    """New mirror instance sharing the parameters of reference module `m`."""
    name = type(m).__name__
    cls = getattr(L, name, None)
    if cls is None or isinstance(m, cls):
        return None
    new = cls.__new__(cls)
    torch.nn.Module.__init__(new)
    new.__dict__.update({k: v for k, v in m.__dict__.items() if not k.startswith('_aivc')})
    return new


def convert(model):
    """Recursively replace reference layer instances by mirrors (in place); returns model."""
    for name, child in list(model.named_children()):
        convert(child)
        sub = _mirror_of(child)
        if sub is not None:
            setattr(model, name, sub)
    return model
