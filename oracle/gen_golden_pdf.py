"""ORACLE (test infrastructure): fixtures for the entropy-model classes, produced by the REFERENCE's own classes
imported from /root/reference/src -- ParametricPdf (laplace / normal, zero_mu, K = 1 and K = 2),
EntropyCoder, PdfParamParameterizer in its mixture modes ('two', 'three', 'gamma'), BallePdfEstim.forward.

    python -m oracle.gen_golden_pdf      # writes tests/golden/pdf_classes.npz (needs /root/reference)
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

REF = os.environ.get('AIVC_REFERENCE_SRC', '/root/reference/src')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, REF)

with contextlib.redirect_stdout(io.StringIO()):
    from layers.entropy_coding.pdf_estimator import ParametricPdf, BallePdfEstim     # noqa: E402
    from layers.entropy_coding.entropy_coder import EntropyCoder                    # noqa: E402
    from layers.misc.misc_layers import PdfParamParameterizer                       # noqa: E402


def main():
    g = torch.Generator().manual_seed(11)
    C, H, W = 6, 9, 13
    fx = {}
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        for mode, nch in (('laplace', 2 * C), ('laplace_two', 5 * C), ('normal_three_gamma', 11 * C)):
            x = torch.randn(1, nch, H, W, generator=g) * 2
            x[0, 0, 0, 0], x[0, -1, 0, 0] = 30., -30.           # exercise the log-var clamps wherever they land
            prm = PdfParamParameterizer(mode, C)(x)
            fx['pp_%s_x' % mode] = x.numpy()
            for k, d in enumerate(prm):
                for name in ('mu', 'sigma', 'gamma', 'weight'):
                    fx['pp_%s_%d_%s' % (mode, k, name)] = d[name].numpy()
            y = torch.round(torch.randn(1, C, H, W, generator=g) * 3)
            fam = 'normal' if 'normal' in mode else 'laplace'
            for zero_mu in (False, True):
                p = ParametricPdf(fam)(y, prm, zero_mu=zero_mu)
                fx['pdf_%s_zero%d_p' % (mode, zero_mu)] = p.numpy()
                fx['pdf_%s_zero%d_rate' % (mode, zero_mu)] = EntropyCoder()(p, y).numpy()
            fx['pdf_%s_y' % mode] = y.numpy()
        torch.manual_seed(5)
        bz = BallePdfEstim(C, pdf_family='')
        z = torch.round(torch.randn(1, C, 4, 5, generator=g) * 2)
        fx['balle_z'] = z.numpy()
        fx['balle_p'] = bz(z).numpy()
        for k, v in bz.state_dict().items():
            fx['balle_sd:' + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, 'pdf_classes.npz'), **fx)
    print('pdf_classes ok:', len(fx), 'arrays')


if __name__ == '__main__':
    main()
