"""GPU parity at the BASELINE.json configurations: the CUDA codec against fixtures the CPU
oracle minted at 416x240 (real BlowingBubbles frame, all intra), 1280x720 (low-delay P: I, P, P) and 1920x1080
(random access: I, P, B; and a 9-frame GOP with B frames three levels deep) for the full-width stand-in (C=128,
Cy=Cz=64) -- oracle/gen_golden_configs.py.

What is asserted, per engine (numbers measured on B200 are in profiles/r02_parity_configs.json):
  fp32 (exact SIMT): at most 1 in 100 000 quantised latent indices differs from the oracle's (measured: 2 of
      2 611 200 at 1080p I/P/B, 56 of 8 878 080 on the 9-frame GOP -- 34 of them in one frame whose reference had a
      one-level tie flip --, 0 of 24 960 at 416x240): fp32 arithmetic in a different summation order than
      MKL-DNN's cannot do better, a pre-rounding value within ~1e-6 of a .5 boundary flips
  bf16x3 (split-bf16 tcgen05, the default and benchmarked engine): at most 1 in 5 000 (measured: 240 of
      2 611 200 = 9.2e-5 at 1080p I/P/B, 1004 of 8 878 080 = 1.1e-4 on the 9-frame GOP with no growth along the
      reference chain, 116 of 1 152 000 at 720p, 0 of 24 960 at 416x240)
  both: never by more than one step; z indices to the same rate; bitstream size within 0.05 %; reconstruction
      within 1 level on the checked subsample; PSNR against the source within 1e-4 dB of the oracle's
      (north_star's tolerance; measured <= 1.6e-6 dB)
  bf16 (plain bf16 operands, the fast mode): <= 3 % of the indices differ; PSNR within 0.01 dB
  every engine: decoder output == encoder reconstruction, bit for bit."""
import pytest
import torch

from tests import parity_cfg

pytestmark = pytest.mark.gpu

CASES = ['bubbles240', 'ldp720', 'ra1080', 'ra1080_gop8']


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU suite needs a CUDA device'
    return torch.device('cuda:0')


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('precision', ['bf16x3', 'fp32'])
def test_exact_engines_match_oracle_indices(case, precision, dev):
    r = parity_cfg.measure(case, precision, dev)
    one_in = 5000 if precision == 'bf16x3' else 100000
    assert r['closed_loop_exact']
    assert r['y_mismatches'] <= max(1, r['y_symbols'] // one_in), r
    assert r['y_max_abs_diff'] <= 1, r
    assert r['z_mismatches'] <= max(1, r['z_symbols'] // one_in), r
    assert abs(r['bytes_delta']) <= max(4, r['oracle_bytes'] // 2000), r
    assert r['max_level_diff_subsampled'] <= 1, r
    assert r['max_abs_psnr_delta_db'] <= 1e-4, r


@pytest.mark.parametrize('case', CASES)
def test_fast_bf16_engine_stays_close(case, dev):
    r = parity_cfg.measure(case, 'bf16', dev)
    assert r['closed_loop_exact']
    assert r['y_index_mismatch_rate'] <= 0.03, r
    assert r['y_max_abs_diff'] <= 2, r
    assert r['max_abs_psnr_delta_db'] <= 1e-2, r
