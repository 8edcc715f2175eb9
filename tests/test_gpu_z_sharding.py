"""Frame-level sharding of ONE GOP with the real FrameCodec (BASELINE.json configs[3]): G processes
share this box's GPU and exchange reconstructions through torch.distributed (gloo here: NCCL refuses two ranks on one
device; the NCCL run over 2/4/8 GPUs is `bench.py --sharding frame`, profiles/r02_frame_sharding_*.json).
Property: the bitstream of a GOP is byte-identical for G = 1, 2, 4 and equals the serial encoder's, and the
sharded decoder reproduces the encoder's reconstruction on every rank."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import hashlib, os, sys
sys.path.insert(0, %r)
import numpy as np
import torch, torch.distributed as dist
from aivc_b200 import models, sharding, gop as G
from aivc_b200.codec import FrameCodec, planes_to_device
from aivc_b200.plan import Config
dist.init_process_group('gloo')
r, n = dist.get_rank(), dist.get_world_size()
dev = torch.device('cuda:0')
torch.cuda.set_device(dev)
h, w = 80, 112
net = models.build_standin(seed=4321, C=32, Cy=16, Cz=16, Csc=16)
gop = G.generate_gop_struct('1_GOP_8')
names = sorted(gop, key=lambda f: int(f.split('_')[1]))
rng = np.random.default_rng(3)
frames = {f: planes_to_device([rng.integers(0, 256, (h, w), dtype=np.uint8), rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8),
                               rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)], dev) for f in names}
for prec in ('bf16x3', 'bf16'):
    codec = FrameCodec(net, h, w, dev, Config(precision=prec))
    stats = {}
    bts, rec = sharding.encode_gop_frame_parallel(codec, frames, gop, stats=stats)
    dec = sharding.decode_gop_frame_parallel(codec, bts, gop)
    for f in names:
        assert all(torch.equal(a, b) for a, b in zip(rec[f], dec[f])), (prec, f)
    serial_b, serial_rec = codec.encode_gop(frames, gop)          # the pipelined single-GPU encoder
    assert bts == serial_b, prec
    for f in names:
        assert all(torch.equal(a, b) for a, b in zip(rec[f], serial_rec[f])), (prec, f)
    if r == 0:
        print('MD5', prec, hashlib.md5(b''.join(bts[f] for f in names)).hexdigest(), 'bcasts', stats.get('bcasts', 0))
dist.destroy_process_group()
'''


def _run(tmp_path, nproc, port):
    script = tmp_path / 'shard_worker.py'
    script.write_text(_WORKER % ROOT)
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(nproc),
                          '--master-addr', '127.0.0.1', '--master-port', str(port), str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    return sorted(l for l in out.stdout.splitlines() if l.startswith('MD5'))


def test_frame_sharded_bitstream_identical_for_1_2_4_ranks(tmp_path):
    md5 = {n: [l.split()[1:3] for l in _run(tmp_path, n, 29650 + n)] for n in (1, 2, 4)}
    assert md5[1] == md5[2] == md5[4], md5
    assert len(md5[1]) == 2
