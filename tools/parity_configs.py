"""Parity of every engine against the oracle on the BASELINE-configuration fixtures -> one JSON document
(profiles/r02_parity_configs.json).  Run on a B200: python tools/parity_configs.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import parity_cfg       # noqa: E402

dev = torch.device('cuda:0')
res = []
for case in ('bubbles240', 'ldp720', 'ra1080', 'ra1080_gop8'):
    for prec in ('fp32', 'bf16x3', 'bf16'):
        r = parity_cfg.measure(case, prec, dev)
        res.append(r)
        print(case, prec, 'y mismatches %d / %d (max |d| %d), z %d / %d, bytes %d vs %d (identical frames %d/%d), planes identical '
              '%d/%d, max level diff %d, max |dPSNR| %.2e dB, closed loop %s'
              % (r['y_mismatches'], r['y_symbols'], r['y_max_abs_diff'], r['z_mismatches'], r['z_symbols'], r['bytes'],
                 r['oracle_bytes'], r['frames_bytes_identical'], r['n_frames'], r['frames_planes_identical'], r['n_frames'],
                 r['max_level_diff_subsampled'], r['max_abs_psnr_delta_db'], r['closed_loop_exact']), flush=True)
        if case == 'ra1080_gop8':        # along the reference chain, in coding order
            print('    per frame (coding order) y mismatches mof/codec:',
                  [(f, fr.get('mof', {}).get('y_mismatches'), fr['codec']['y_mismatches']) for f, fr in r['frames'].items()], flush=True)
json.dump(res, open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/r02_parity_configs.json', 'w'), indent=1)
