"""GOP schedules (I/P/B types, references, coding order).

Same dict layout and naming as the reference (func_util/GOP_structure.py:27-137,
199-221): {'frame_<display idx>': {'type', 'prev_ref', 'next_ref', 'coding_order'}} so the
bitstream's GOP header (header.py:129-210) keeps meaning the same schedule.  Built
iteratively here (the reference recurses); ``levels`` adds the dependency-level view
the multi-GPU scheduler uses.
"""
FRAME_I, FRAME_P, FRAME_B = 0, 1, 2


def _entry(t, prev, nxt, order):
    name = lambda i: None if i is None else 'frame_%d' % i
    return {'type': t, 'prev_ref': name(prev), 'next_ref': name(nxt), 'coding_order': order}


def _ra(gop_size, base=0, order0=0, with_intra=True):
    """One random-access GOP: I(base) [optional], P(base+gop_size), then the B pyramid in
    depth-first (left before right) order, which is the reference's coding order."""
    gop, order = {}, order0
    if with_intra:
        gop['frame_%d' % base] = _entry(FRAME_I, None, None, order)
    order += 1
    gop['frame_%d' % (base + gop_size)] = _entry(FRAME_P, base, None, order)
    order += 1
    half = gop_size // 2
    stack = [(base + half, half)] if half > 0 else []
    while stack:
        idx, n = stack.pop()
        gop['frame_%d' % idx] = _entry(FRAME_B, idx - n, idx + n, order)
        order += 1
        n2 = n // 2
        if n2:
            stack.append((idx + n2, n2))      # right pushed first -> left is coded first
            stack.append((idx - n2, n2))
    return gop


def generate_gop_struct(name):
    """'1_GOP_0' (all intra), 'LDP_<n>' (I + n P), '<k>_GOP_<n>' (k chained RA GOPs of size n)."""
    toks = name.split('_')
    if name == '1_GOP_0':
        return {'frame_0': _entry(FRAME_I, None, None, 0)}
    if 'LDP' in toks:
        n = int(toks[-1])
        gop = {'frame_0': _entry(FRAME_I, None, None, 0)}
        for i in range(1, n + 1):
            gop['frame_%d' % i] = _entry(FRAME_P, i - 1, None, i)
        return gop
    k, n = int(toks[0]), int(toks[-1])
    gop = _ra(n)
    for i in range(1, k):
        gop.update(_ra(n, base=i * n, order0=i * n, with_intra=False))
    return gop


def coding_order(gop):
    return sorted(gop, key=lambda f: gop[f]['coding_order'])


def levels(gop):
    """Frames grouped by dependency depth: every frame of level L only needs frames of
    levels < L, so one level can be coded in parallel across GPUs."""
    depth = {}
    for f in coding_order(gop):
        refs = [r for r in (gop[f]['prev_ref'], gop[f]['next_ref']) if r is not None]
        depth[f] = 1 + max((depth[r] for r in refs), default=-1)
    out = [[] for _ in range(max(depth.values()) + 1)]
    for f in coding_order(gop):
        out[depth[f]].append(f)
    return out
