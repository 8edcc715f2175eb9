"""CPU suite: bitstream container known answers (bytes produced by the reference's own
header.py / cat_binary_files.py for the same inputs, captured when the fixtures were made),
module-path aliases for un-pickling, the torchac shim, and 2-rank GOP sharding over gloo."""
import os
import pickle
import subprocess
import sys

import numpy as np
import torch

from aivc_b200 import container as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# reference output of write_gop_header + cat_one_gop + cat_one_video for frames 3,4,5 =
# b'abc'*5, b'', b'xyz'; '1_GOP_2'; idx_rate 0.5; dims (1080,1920)/(68,120)/(17,30)
GOP_KAT = '0000010002080000000f616263616263616263616263616263000000000000000378797a'
VIDEO_KAT = ('04380780004400780011001e00010003000500000024' + GOP_KAT)


def test_container_known_answers():
    g = K.pack_gop('1_GOP_2', [b'abc' * 5, b'', b'xyz'], 0.5)
    assert g.hex() == GOP_KAT
    v = K.pack_video((1080, 1920), (68, 120), (17, 30), [g], 3, 5)
    assert v.hex() == VIDEO_KAT
    dims, gops, first, last = K.unpack_video(v)
    assert dims == {'x': (1080, 1920), 'y': (68, 120), 'z': (17, 30), 'x_uv': (540, 960)}
    assert (first, last) == (3, 5) and gops == [g]
    name, rate, frames = K.unpack_gop(gops[0])
    assert (name, rate, frames) == ('1_GOP_2', 0.5, [b'abc' * 5, b'', b'xyz'])
    assert K.parse_gop_header(K.gop_header('LDP_8')) == ('LDP_8', 0.0)


def test_install_aliases_allow_unpickling_reference_paths():
    from aivc_b200 import compat
    import aivc_b200.layers as L
    compat.install()
    import layers.misc.custom_conv_layers as ref_path
    assert ref_path.ChengResBlock is L.ChengResBlock
    import models
    assert hasattr(models, 'FullNet')
    m = L.ChengResBlock(8, 'down')
    # a pickle that names the REFERENCE module path must resolve to the mirror
    saved = {}
    for path, names in compat._MODULE_MAP.items():
        for n in names:
            saved[n] = getattr(L, n).__module__
            getattr(L, n).__module__ = path
    try:
        blob = pickle.dumps(m)
    finally:
        for n, mod in saved.items():
            getattr(L, n).__module__ = mod
    assert b'layers.misc.custom_conv_layers' in blob and b'aivc_b200.layers' not in blob
    m2 = pickle.loads(blob)
    assert type(m2) is L.ChengResBlock
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    assert hasattr(torch, 'set_deterministic')


def test_torchac_shim_matches_oracle_coder():
    from aivc_b200 import compat
    from oracle import torchac_shim as O
    compat.install()
    import torchac
    g = torch.Generator().manual_seed(0)
    cdf = torch.sort(torch.rand(200, 514, generator=g), dim=1)[0]
    cdf[:, 0], cdf[:, -1] = 0, 1
    sym = torch.randint(0, 512, (200,), generator=g).to(torch.int16)
    a = torchac.encode_float_cdf(cdf, sym, check_input_bounds=True)
    assert a == O.encode_float_cdf(cdf, sym)
    assert torch.equal(torchac.decode_float_cdf(cdf, a), sym)


def test_convert_swaps_reference_named_layers():
    from aivc_b200 import compat
    import aivc_b200.layers as L

    class CustomConvLayer(torch.nn.Module):      # stands for the reference class of that name
        def __init__(self):
            super().__init__()
            self.layers = torch.nn.Sequential(torch.nn.ReplicationPad2d(1), torch.nn.Conv2d(4, 4, 3))

    holder = torch.nn.Sequential(CustomConvLayer(), torch.nn.Sequential(CustomConvLayer()))
    w = holder[0].layers[1].weight
    compat.convert(holder)
    assert type(holder[0]) is L.CustomConvLayer and type(holder[1][0]) is L.CustomConvLayer
    assert holder[0].layers[1].weight is w                    # shared, not copied


_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from aivc_b200 import sharding
dist.init_process_group('gloo')
r, n = dist.get_rank(), dist.get_world_size()
mine = sharding.gops_of_rank(7, r, n)
payload = {g: bytes([g]) * (g + 1) for g in mine}
allb = sharding.gather_gop_bytes(payload, 7)
if r == 0:
    assert [len(b) for b in allb] == [i + 1 for i in range(7)], allb
    assert all(b == bytes([i]) * (i + 1) for i, b in enumerate(allb))
    print('OK', mine)
dist.destroy_process_group()
'''


_WORKER2 = r'''
import os, sys, hashlib
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from aivc_b200 import sharding, gop as G
dist.init_process_group('gloo')
r, n = dist.get_rank(), dist.get_world_size()
class FakeCodec:
    """Deterministic stand-in: the 'reconstruction' mixes the frame with its references, so a
    wrong or missing reference exchange changes every later frame."""
    h, w, device = 6, 8, torch.device('cpu')          # 48 luma + 2 x 12 chroma samples
    def __init__(self):
        self.store = {}
    def encode_frame(self, planes, t, prev, nxt):
        rec = tuple((p.to(torch.int32) * 3 + (0 if prev is None else prev[i].to(torch.int32))
                     + 2 * (0 if nxt is None else nxt[i].to(torch.int32)) + t).remainder(251).to(torch.uint8)
                    for i, p in enumerate(planes))
        key = hashlib.md5(b''.join(x.numpy().tobytes() for x in rec)).digest()
        return key, rec
    def decode_frame(self, b, t, prev, nxt):
        return self.store[b]
gop = G.generate_gop_struct('1_GOP_8')
g = torch.Generator().manual_seed(0)
sizes = [48, 12, 12]
frames = {f: tuple(torch.randint(0, 256, (s,), dtype=torch.uint8, generator=g) for s in sizes) for f in sorted(gop)}
codec = FakeCodec()
stats = {}
bts, rec = sharding.encode_gop_frame_parallel(codec, frames, gop, stats=stats)
# serial reference on every rank
ref, ref_b = {}, {}
for f in G.coding_order(gop):
    e = gop[f]
    ref_b[f], ref[f] = codec.encode_frame(frames[f], e['type'], ref.get(e['prev_ref']), ref.get(e['next_ref']))
    codec.store[ref_b[f]] = ref[f]
assert all(all(torch.equal(a, b) for a, b in zip(rec[f], ref[f])) for f in gop)
assert bts == ref_b                                   # the whole GOP's bytes on EVERY rank, identical to the serial run
assert stats['bcasts'] == len(gop)
dec = sharding.decode_gop_frame_parallel(codec, bts, gop)
assert all(all(torch.equal(a, b) for a, b in zip(dec[f], ref[f])) for f in gop)
# one rank alone (no process group use): same bytes
solo_b, _ = sharding.encode_gop_frame_parallel(codec, frames, gop, rank=0, world=1)
assert solo_b == ref_b
if r == 0:
    print('OK levels', [len(l) for l in G.levels(gop)])
dist.destroy_process_group()
'''


def test_frame_level_sharding_two_ranks_gloo(tmp_path):
    script = tmp_path / 'w2.py'
    script.write_text(_WORKER2 % ROOT)
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', '29543', str(script)],
                         capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'OK levels [1, 1, 1, 2, 4]' in out.stdout


def test_gop_sharding_two_ranks_gloo(tmp_path):
    script = tmp_path / 'w.py'
    script.write_text(_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29541')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', '29541', str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'OK [0, 2, 4, 6]' in out.stdout


def _run_py(code, cwd, extra_env=None):
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, 'compat'))
    env.update(extra_env or {})
    return subprocess.run([sys.executable, '-c', code], cwd=cwd, env=env, capture_output=True, text=True, timeout=300)


def test_sitecustomize_is_subprocess_safe_and_loads_whole_module_pickles(tmp_path):
    """compat/sitecustomize.py on PYTHONPATH (how aivc.py's encode.py / decode.py subprocesses get the shims,
    src/aivc.py:117-139): a fresh interpreter resolves the reference's module paths to the mirrors, has torchac and
    torch.set_deterministic, and torch.load()s a whole-module pickle that names reference paths
    (model_management.py:347) with the default arguments."""
    from aivc_b200 import compat, models
    import aivc_b200.layers as L
    compat.install()              # (pickle checks that the renamed module paths resolve to these very classes)
    net = models.build_standin(seed=3, C=16, Cy=8, Cz=8, Csc=8)
    saved = {}
    classes = [(path, n) for path, names in compat._MODULE_MAP.items() for n in names] + \
              [('models', n) for n in ('FullNet', 'ConditionalNet', '_Wrap', 'MotionCompensation')]
    for path, n in classes:
        cls = getattr(L, n, None) or getattr(models, n)
        saved[cls] = cls.__module__
        cls.__module__ = path
    try:
        torch.save(net, str(tmp_path / '0_model.pt'))
    finally:
        for cls, mod in saved.items():
            cls.__module__ = mod
    code = ('import torch, torchac, models, layers.misc.misc_layers as ml, layers.entropy_coding.entropy_coder as ec\n'
            'from layers.entropy_coding.pdf_estimator import ParametricPdf, BallePdfEstim\n'
            'assert hasattr(torch, "set_deterministic") and hasattr(ml, "View") and hasattr(ml, "LowerBound")\n'
            '# the unsafe whole-module unpickling is scoped to the reference\'s own call site (model_management.py:347)\n'
            'import types, pickle\n'
            'mm = types.ModuleType("model_mngt.model_management"); mm.torch = torch\n'
            'exec("def load_model(p):\\n    return torch.load(p, map_location=\'cpu\')", mm.__dict__)\n'
            'try:\n'
            '    torch.load("0_model.pt", map_location="cpu"); raise SystemExit("plain torch.load must stay weights_only")\n'
            'except pickle.UnpicklingError:\n'
            '    pass\n'
            'm = mm.load_model("0_model.pt")\n'
            'cn = m.codec_net.codec_net\n'
            'assert type(m).__module__ == "aivc_b200.models" and type(cn.g_a[0]).__module__ == "aivc_b200.layers"\n'
            'assert hasattr(m, "GOP_forward") and cn.nb_ft_y == 8 and isinstance(cn.pdf_y, ParametricPdf)\n'
            'print("LOADED", sum(p.numel() for p in m.parameters()))\n')
    out = _run_py(code, str(tmp_path))
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'LOADED %d' % sum(p.numel() for p in net.parameters()) in out.stdout
    off = _run_py('import torchac', str(tmp_path), {'AIVC_B200_NO_COMPAT': '1'})
    assert off.returncode != 0                     # (the switch really switches it off; torchac is not installed)


def test_sitecustomize_makes_the_reference_modules_importable():
    """real_life.bitstream, real_life.decode and model_mngt.model_management do not import on this
    torch without torchac; with compat/ on PYTHONPATH they do, from the reference's own src/ (skipped where the
    reference tree is absent, e.g. on the GPU box), and the reference's ArithmeticCoder round-trips a latent through
    the shimmed torchac."""
    src = '/root/reference/src'
    if not os.path.isdir(src):
        import pytest
        pytest.skip('reference tree not present')
    code = ('import io, contextlib, torch\n'
            'with contextlib.redirect_stdout(io.StringIO()):\n'
            '    import real_life.bitstream as B, real_life.decode as D, model_mngt.model_management as MM\n'
            '    from func_util.cluster_mngt import seed_all\n'
            '    seed_all(seed=666)\n'
            'import layers.misc.custom_conv_layers as C\n'
            'assert C.CustomConvLayer.__module__ == "aivc_b200.layers"\n'
            'print("IMPORTED", hasattr(D, "Decoder"), hasattr(MM, "load_model"), hasattr(B, "ArithmeticCoder"))\n')
    out = _run_py(code, src)
    assert out.returncode == 0, out.stderr[-3000:]
    assert 'IMPORTED True True True' in out.stdout
