"""Integer-exact synthetic 4:2:0 clips for parity fixtures: every operation is integer numpy, so the frames are
bit-identical on any machine (the float box filters of bench.synth_gop are not guaranteed to be).
Content as SURVEY.md 8d prescribes: a low-pass noise texture translated by (2t, t) pixels per frame plus 5 %
fresh noise, chroma = 2x2 mean of luma (V mirrored), 8 bit."""
import numpy as np


def _box(a, r):
    """(2r+1) x (2r+1) box SUM of an int64 image with edge replication, exact."""
    k = 2 * r + 1
    p = np.pad(a, r, mode='edge')
    c = np.cumsum(np.pad(p, ((1, 0), (0, 0))), axis=0)
    p = c[k:] - c[:-k]
    c = np.cumsum(np.pad(p, ((0, 0), (1, 0))), axis=1)
    return c[:, k:] - c[:, :-k]


def clip(seed, n_frames, h, w):
    """-> list of (y [h,w], u [hc,wc], v [hc,wc]) uint8 arrays."""
    rng = np.random.default_rng(seed)
    hh, ww = h + n_frames + 8, w + 2 * n_frames + 8
    base = rng.integers(0, 256, (hh, ww), dtype=np.int64)
    for _ in range(2):
        base = _box(base, 8) // 289
    lo, hi = int(base.min()), int(base.max())
    base = (base - lo) * 255 // max(hi - lo, 1)
    hc, wc = (h + 1) // 2, (w + 1) // 2
    out = []
    for t in range(n_frames):
        y = base[t:t + h, 2 * t:2 * t + w]
        y = (95 * y + 5 * rng.integers(0, 256, (h, w), dtype=np.int64) + 50) // 100
        yp = np.pad(y, ((0, 2 * hc - h), (0, 2 * wc - w)), mode='edge')
        u = (yp.reshape(hc, 2, wc, 2).sum(axis=(1, 3)) + 2) // 4
        out.append((y.astype(np.uint8), u.astype(np.uint8), (255 - u).astype(np.uint8)))
    return out
