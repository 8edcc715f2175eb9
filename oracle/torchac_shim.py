"""ORACLE (test infrastructure only).

A module with torchac's three public entry points used by AIVC
(``encode_float_cdf``, ``decode_float_cdf`` -- src/real_life/bitstream.py:281,454,482),
backed by ``oracle/torchac_ref.c``.  Installed as ``sys.modules['torchac']`` by
``oracle/gen_golden.py`` so that the reference's own ``real_life.bitstream`` /
``real_life.decode`` import and run in this container.
"""
import numpy as np
import torch

from . import codec_ref as C
from .nn_ref import cdf_float_to_int


def _to_int(cdf_float, needs_normalization):
    if not needs_normalization:
        return ((cdf_float * 65536.0).round().to(torch.int64) & 0xFFFF)
    return cdf_float_to_int(cdf_float)


def encode_float_cdf(cdf_float, sym, needs_normalization=True, check_input_bounds=False):
    if check_input_bounds:
        if cdf_float.min() < 0 or cdf_float.max() > 1:
            raise ValueError('cdf_float out of [0, 1]')
        if sym.max() >= cdf_float.shape[-1] - 1 or sym.min() < 0:
            raise ValueError('symbol out of range')
    lp = cdf_float.shape[-1]
    table = _to_int(cdf_float, needs_normalization).numpy().astype(np.uint16).reshape(-1, lp)
    return C.rc_encode_table(table, sym.reshape(-1).numpy().astype(np.int16))


def decode_float_cdf(cdf_float, byte_stream, needs_normalization=True):
    lp = cdf_float.shape[-1]
    table = _to_int(cdf_float, needs_normalization).numpy().astype(np.uint16).reshape(-1, lp)
    n = table.shape[0]
    sym = C.rc_decode_table(table, byte_stream, n)
    return torch.from_numpy(sym.copy()).view(cdf_float.shape[:-1])
