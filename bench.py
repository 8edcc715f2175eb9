#!/usr/bin/env python
"""Headline benchmark: 1080p frames/s, encode + decode, synthetic YUV420, stand-in AIVC model.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

One step = one random-access GOP ('1_GOP_32': I0, P32 and 31 B frames = 33 frames) of
1920x1080 pushed through encoder and decoder (BASELINE.json configs[2]).  With N GPUs every
rank codes its own GOPs (GOPs are independent units: SURVEY.md 8e) -- weak scaling, no
data-path collective.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 1080, 1920
GOP_NAME = '1_GOP_32'
MODEL = dict(seed=1234, C=128, Cy=64, Cz=64, Csc=64)
METRIC = '1080p frames/sec encode+decode'


def synth_gop(seed, n_frames, h=H, w=W):
    """Seeded synthetic 4:2:0 clip: low-pass noise texture translated by (2t, t)
    pixels per frame plus 5% fresh noise, 8-bit."""
    rng = np.random.default_rng(seed)
    import torch
    import torch.nn.functional as F
    base = torch.from_numpy(rng.random((1, 1, h + 2 * n_frames + 32, w + 4 * n_frames + 32), dtype=np.float32))
    k = torch.ones(1, 1, 1, 17) / 17.0
    for _ in range(2):
        base = F.conv2d(F.conv2d(base, k, padding=(0, 8)), k.transpose(2, 3), padding=(8, 0))
    base = (base - base.min()) / (base.max() - base.min())
    frames = []
    for t in range(n_frames):
        y = base[0, 0, t:t + h, 2 * t:2 * t + w].numpy()
        y = np.clip(0.95 * y + 0.05 * rng.random((h, w), dtype=np.float32), 0, 1)
        u = y.reshape(h // 2, 2, w // 2, 2).mean(axis=(1, 3))
        yy = np.rint(y * 255).astype(np.uint8)
        uu = np.rint(u * 255).astype(np.uint8)
        frames.append((yy, uu, (255 - uu).astype(np.uint8)))
    return frames


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '200'],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(nme)
        os.unlink(self.f.name)
        if sm:
            out = {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                   'samples': len(sm)}
        return out


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops_sustained', 1362.6), d.get('hbm_gbs', 6547.8), 'measured (MEASURED_PEAKS.json, sustained bf16)'
    return 1400.0, 6650.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------ reference arm
def cpu_reference_sample(budget_s, threads=None):
    """Times the oracle (CPU restatement of the reference's algorithm, fp32 torch on host cores)
    on a bounded sample: encode + decode of the GOP's I frame, on a full-width crop of the 1080p
    frame whose height is fitted to `budget_s`.  Conv cost is linear in area, so
    frames/s at 1080p = (crop_rows / 1080) / seconds."""
    import torch
    from aivc_b200 import models
    from oracle import codec_ref as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    net = models.build_standin(**MODEL)
    tables = O.Tables(net)

    def run(rows):
        fr = synth_gop(0, 1, rows, W)[0]
        yuv = {k: torch.from_numpy(p.astype(np.float32) / 255.)[None, None] for k, p in zip('yuv', fr)}
        z = O.zero_yuv(rows, W)
        t0 = time.time()
        data, rec, _ = O.encode_frame(net, tables, yuv, z, z, 0)
        dec, _ = O.decode_frame(net, tables, data, z, z, 0, rows, W)
        dt = time.time() - t0
        assert all(torch.equal(rec[k], dec[k]) for k in 'yuv')
        return dt

    t_probe = run(64)                               # calibration crop (also the warm-up)
    rows = int(min(H, max(64, (budget_s / max(t_probe, 1e-3)) * 64)) // 8 * 8)
    dt = run(rows)
    fps = (rows / H) / dt
    return fps, threads, ('I-frame encode+decode of a %dx%d crop of the 1080p frame (%.1f s), scaled by '
                          'area to 1080p; inter frames cost ~2.9x more FLOPs, so this flatters the CPU'
                          % (W, rows, dt))


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n_steps = args.steps + args.warmup
    budget = max(2.0, min(20.0, 150.0 / max(n_steps, 1)))
    vals, sample, threads = [], '', None
    for i in range(n_steps):
        fps, threads, sample = cpu_reference_sample(budget)
        if i >= args.warmup:
            vals.append(fps)
    v = float(np.mean(vals))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'frames/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000.0 / v * 33, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.gpus),
        'cpu_baseline': {'value': v, 'unit': 'frames/s', 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {'workload': 'Random Access %s (33 frames/GOP), synthetic 1920x1080 YUV420, stand-in AIVC model '
                        '(MOFNet+CodecNet, C=128, Cy=Cz=64, seed 1234); 1 GOP per GPU per step' % GOP_NAME,
            'frames_per_step_per_gpu': 33, 'gop': GOP_NAME, 'resolution': '1920x1080',
            'sharding': 'one GOP per rank, no data-path collective',
            'l2': 'working set per step (activations > 2 GB, 102 MB of frames) exceeds the 126 MB L2'}


# ------------------------------------------------------------------------------ CUDA arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from aivc_b200 import models, gop as G, _lib
    from aivc_b200.codec import FrameCodec
    from aivc_b200.plan import Config

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    L = _lib.lib()

    net = models.build_standin(**MODEL)
    gop = G.generate_gop_struct(GOP_NAME)
    names = sorted(gop, key=lambda f: int(f.split('_')[1]))
    codec = FrameCodec(net, H, W, dev, Config(precision=args.precision))
    clip = synth_gop(100 + rank, len(names))
    host = {f: tuple(torch.from_numpy(p.reshape(-1)).pin_memory() for p in clip[i]) for i, f in enumerate(names)}
    resident = {f: tuple(p.to(dev) for p in host[f]) for f in names}
    out_host = {f: tuple(torch.empty_like(p) for p in host[f]) for f in names}
    frame_bytes = sum(p.numel() for p in host[names[0]])
    state = {}

    def step(e2e):
        if e2e:
            frames = {f: tuple(p.to(dev, non_blocking=True) for p in host[f]) for f in names}
        else:
            frames = resident
        t0 = time.perf_counter()
        bts, rec = codec.encode_gop(frames, gop)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        dec = codec.decode_gop(bts, gop)
        torch.cuda.synchronize()
        state['enc_s'] = state.get('enc_s', 0.0) + (t1 - t0)
        state['dec_s'] = state.get('dec_s', 0.0) + (time.perf_counter() - t1)
        if e2e:
            for f in names:
                for d, s in zip(out_host[f], dec[f]):
                    d.copy_(s, non_blocking=True)
        state['bts'], state['rec'], state['dec'] = bts, rec, dec

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e, profile):
        barrier()
        l0 = L.aivc_launch_count()
        if profile:
            L.aivc_profile_enable(1)
        # per-stage timing needs kernels one at a time: no second stream next to the timed stages
        # (the library likewise drops its two-lane execution while profiling)
        overlap = codec.mof.overlap_shortcut
        codec.mof.overlap_shortcut = codec.codec.overlap_shortcut = overlap and not profile
        sampler = ClockSampler(local) if rank == 0 else None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            step(e2e)
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        clocks = sampler.stop() if sampler else None
        prof = None
        if profile:
            out = (C.c_double * 6)()
            _lib.check(L.aivc_profile_read(out))
            ncls = len(_lib.KERNEL_CLASSES)
            cls = (C.c_double * (3 * ncls))()
            _lib.check(L.aivc_profile_read_classes(cls, ncls))
            prof = list(out) + list(cls)
            if args.stage_csv and rank == 0:
                _lib.check(L.aivc_profile_dump(args.stage_csv.encode()))
            L.aivc_profile_enable(0)
        codec.mof.overlap_shortcut = codec.codec.overlap_shortcut = overlap
        launches = L.aivc_launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks, prof, launches

    for _ in range(args.warmup):
        step(False)
    state['enc_s'] = state['dec_s'] = 0.0
    ms, clocks, _, launches = timed(False, False)
    enc_ms, dec_ms = 1e3 * state['enc_s'] / args.steps, 1e3 * state['dec_s'] / args.steps
    # roofline leg: the same K steps again with a CUDA-event pair around every convolution stage
    # (the event records sit between kernels, which costs ~4% of step time, so the headline
    # `value` above is taken without them)
    ms_prof, _, prof, _ = timed(False, True)
    # closed loop must hold on the benchmarked data (decoder == encoder reconstruction)
    for f in names:
        for x, y in zip(state['rec'][f], state['dec'][f]):
            assert torch.equal(x, y), 'closed loop broken on ' + f
    total_bytes = sum(len(b) for b in state['bts'].values())
    step(True)
    ms_e2e, _, _, _ = timed(True, False)

    frames_per_step = len(names) * world
    value = frames_per_step * args.steps / (ms / 1000.0)
    e2e = frames_per_step * args.steps / (ms_e2e / 1000.0)
    if rank == 0:
        peak_tf, peak_bw, peak_src = peaks()
        tc_ms, tc_fl, tc_n, si_ms, si_fl, si_n = prof[:6]
        ach_all = tc_fl / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        # per kernel: algorithmic FLOPs of its launches / their CUDA-event durations
        by_kernel, all_ms = {}, sum(prof[6 + 3 * k] for k in range(len(_lib.KERNEL_CLASSES)))
        for k, name in enumerate(_lib.KERNEL_CLASSES):
            k_ms, k_fl, k_n = prof[6 + 3 * k: 9 + 3 * k]
            if k_n:
                by_kernel[name] = {'launches_per_step': k_n / args.steps, 'avg_launch_us': 1e3 * k_ms / k_n,
                                   'tflops': k_fl / (k_ms * 1e-3) / 1e12, 'share_of_stage_time': k_ms / all_ms}
        dom = max((n for n in by_kernel if by_kernel[n]['tflops'] > 0), key=lambda n: by_kernel[n]['share_of_stage_time'])
        ach = by_kernel[dom]['tflops']
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'r01_traffic.json')     # dram bytes per launch from the ncu --set full captures
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(dom)
        line = {
            'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': {'bf16': 'bf16', 'bf16x3': 'bf16x3 (split bf16 operands, fp32 accumulate)', 'fp32': 'f32'}[args.precision],
            'data': 'synthetic', 'config': workload_config(world), 'clocks': clocks,
            'e2e': {'value': e2e, 'unit': 'frames/s',
                    'h2d_bytes_per_step': frame_bytes * len(names), 'd2h_bytes_per_step': frame_bytes * len(names),
                    'note': 'frames start in pinned host memory and decoded planes return to pinned host memory '
                            'every step; the bitstream (bytes) is produced/consumed on the host in both modes'},
            'gpu_launches': int(launches),
            'roofline': {
                'bound': 'tensor', 'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach / peak_tf,
                'traffic': traffic, 'peak_source': peak_src,
                'kernel': dom + ' (the kernel with the largest share of GPU time)',
                'how': 'algorithmic FLOPs of this kernel\'s launches in K timed steps / sum of their CUDA-event '
                       'durations on the launching stream (second pass of the same K steps, an event pair '
                       'around every convolution stage); `traffic` = DRAM bytes of one representative launch '
                       '(ncu --set full, profiles/)',
                'launches_per_step': by_kernel[dom]['launches_per_step'],
                'avg_launch_us': by_kernel[dom]['avg_launch_us'],
                'share_of_stage_time': by_kernel[dom]['share_of_stage_time'],
                'by_kernel': by_kernel,
                'all_tensor_stages': {'tflops': ach_all, 'frac': ach_all / peak_tf, 'stages': int(tc_n)},
                'tc_ms_per_step': tc_ms / args.steps, 'tc_share_of_step': tc_ms / ms_prof,
                'profiled_ms_per_step': ms_prof / args.steps,
                'simt_ms_per_step': si_ms / args.steps, 'simt_tflops': si_fl / max(si_ms, 1e-9) / 1e9,
            },
            'bitstream_bytes_per_gop': total_bytes, 'encode_ms_per_gop': enc_ms, 'decode_ms_per_gop': dec_ms,
            'encode_fps': 33e3 / enc_ms, 'decode_fps': 33e3 / dec_ms,
            'encode_decode_closed_loop': True,
        }
        if world == 1 and not args.no_cpu_baseline:
            fps, threads, sample = cpu_reference_sample(15.0)
            line['cpu_baseline'] = {'value': fps, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
                                    'sample': sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default='bf16x3', choices=['bf16x3', 'bf16', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--stage-csv', default='', help='dump per-stage CUDA-event timings of the timed region')
    args = ap.parse_args()
    import __graft_entry__ as g
    if int(os.environ.get('LOCAL_RANK', '0')) == 0:
        if not os.path.exists(os.path.join(ROOT, 'aivc_b200', 'libaivc_b200.so')):
            g.build()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
