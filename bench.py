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
# frames of each type (I, P, B) in one GOP, and the 3-frame GOP the CPU arm samples (one frame of every type present)
TYPE_COUNTS, SAMPLE_GOP = {0: 1, 1: 1, 2: 31}, '1_GOP_2'
WORKLOAD = 'ra1080'


def set_workload(name, seed=None):
    """BASELINE.json configs: 'ra1080' = configs[2] (the headline; configs[3] is its --sharding frame form, configs[4]
    its --model-seed sweep), 'ldp720' = configs[1] (low-delay P, 1280x720, GOP 8: I + 8 P, stand-in seed 4)."""
    global H, W, GOP_NAME, MODEL, METRIC, TYPE_COUNTS, SAMPLE_GOP, WORKLOAD
    WORKLOAD = name
    if name == 'ldp720':
        H, W, GOP_NAME, METRIC = 720, 1280, 'LDP_8', '720p frames/sec encode+decode'
        MODEL = dict(MODEL, seed=4)
        TYPE_COUNTS, SAMPLE_GOP = {0: 1, 1: 8, 2: 0}, 'LDP_2'
    if seed is not None:
        MODEL = dict(MODEL, seed=seed)


def synth_gop(seed, n_frames, h=None, w=None):
    """Seeded synthetic 4:2:0 clip: low-pass noise texture translated by (2t, t)
    pixels per frame plus 5% fresh noise, 8-bit."""
    h, w = h or H, w or W                      # (the workload's size at call time)
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
_CPU_MODEL = {}


def cpu_reference_sample(budget_s, threads=None):
    """Times the oracle (CPU restatement of the reference's algorithm, fp32 torch on host cores) on a bounded
    sample of the benchmarked workload: ONE frame of every type -- I, P and B ('1_GOP_2') -- encoded AND decoded on a
    full-width crop of the 1080p clip whose height is fitted to `budget_s`.  The GOP figure is composed by
    frame-type counts (1_GOP_32 = 1 I + 1 P + 31 B) and scaled by area (the convolutions, >95 % of the time, are
    linear in area):  frames/s at 1080p = 33 / ((t_I + t_P + 31 t_B) * 1080 / rows)."""
    import torch
    from aivc_b200 import models, gop as G
    from oracle import codec_ref as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    key = repr(sorted(MODEL.items()))
    if key not in _CPU_MODEL:                          # (built once per process, not once per step)
        net = models.build_standin(**MODEL)
        _CPU_MODEL[key] = (net, O.Tables(net))
    net, tables = _CPU_MODEL[key]
    gop = G.generate_gop_struct(SAMPLE_GOP)
    order = sorted(gop, key=lambda f: gop[f]['coding_order'])
    n_of = {t: sum(1 for f in gop if gop[f]['type'] == t) for t in (0, 1, 2)}

    def run(rows):
        clip = synth_gop(0, 3, rows, W)
        yuv = {'frame_%d' % i: {k: torch.from_numpy(p.astype(np.float32) / 255.)[None, None] for k, p in zip('yuv', fr)}
               for i, fr in enumerate(clip)}
        t = {0: 0.0, 1: 0.0, 2: 0.0}
        rec, dec, bts = {}, {}, {}
        for f in order:                                  # encoder (closed loop), frame by frame
            ft = gop[f]['type']
            prev = rec[gop[f]['prev_ref']] if ft != 0 else O.zero_yuv(rows, W)
            nxt = rec[gop[f]['next_ref']] if ft == 2 else O.zero_yuv(rows, W)
            t0 = time.time()
            bts[f], rec[f], _ = O.encode_frame(net, tables, yuv[f], prev, nxt, ft)
            t[ft] += time.time() - t0
        for f in order:                                  # decoder
            ft = gop[f]['type']
            prev = dec[gop[f]['prev_ref']] if ft != 0 else O.zero_yuv(rows, W)
            nxt = dec[gop[f]['next_ref']] if ft == 2 else O.zero_yuv(rows, W)
            t0 = time.time()
            dec[f], _ = O.decode_frame(net, tables, bts[f], prev, nxt, ft, rows, W)
            t[ft] += time.time() - t0
            assert all(torch.equal(rec[f][k], dec[f][k]) for k in 'yuv')
        return t

    probe_rows = 32
    tp = run(probe_rows)                            # calibration crop (also the warm-up)
    per_row = sum(tp.values()) / probe_rows
    rows = int(min(H, max(32, budget_s / max(per_row, 1e-4))) // 16 * 16)
    t = run(rows)
    per = {k: t[k] / max(n_of[k], 1) for k in t}                 # seconds per frame of each type
    n_frames = sum(TYPE_COUNTS.values())
    gop_s = sum(TYPE_COUNTS[k] * per[k] for k in per) * (H / rows)
    fps = n_frames / gop_s
    return fps, threads, ('one GOP %s (frame types I/P/B: %d/%d/%d) encode+decode of a %dx%d crop of the %dx%d clip: '
                          '%.2f / %.2f / %.2f s per I / P / B frame; GOP %s of %d frames = %d t_I + %d t_P + %d t_B = %.0f s '
                          'after scaling by area (x %.2f)'
                          % (SAMPLE_GOP, n_of[0], n_of[1], n_of[2], W, rows, W, H, per[0], per[1], per[2], GOP_NAME, n_frames,
                             TYPE_COUNTS[0], TYPE_COUNTS[1], TYPE_COUNTS[2], gop_s, H / rows)), gop_s


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n_steps = args.steps + args.warmup
    budget = max(3.0, min(40.0, 170.0 / max(n_steps, 1)))
    vals, gops, walls, sample, threads = [], [], [], '', None
    for i in range(n_steps):
        t0 = time.time()
        fps, threads, sample, gop_s = cpu_reference_sample(budget)
        if i >= args.warmup:
            vals.append(fps)
            gops.append(gop_s)
            walls.append(time.time() - t0)
    v = float(np.mean(vals))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'frames/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup,
        # a step of THIS arm is a bounded sample (see cpu_baseline.sample), and ms_per_step is what such a step took here;
        # the whole-GOP step the CUDA arm times would take full_step_ms_extrapolated on these cores
        'ms_per_step': 1e3 * float(np.mean(walls)), 'step_is_bounded_sample': True,
        'full_step_ms_extrapolated': 1e3 * float(np.mean(gops)),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.gpus),
        'note': 'value = frames of one GOP / (seconds per I, P, B frame measured in the sample, composed by the GOP\'s '
                'frame-type counts and scaled by crop area); the convolutions (>95 % of the time) are linear in area',
        'cpu_baseline': {'value': v, 'unit': 'frames/s', 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def workload_config(n_gpus, sharding='gop'):
    n_frames = sum(TYPE_COUNTS.values())
    return {'workload': '%s %s (%d frames/GOP), synthetic %dx%d YUV420, stand-in AIVC model '
                        '(MOFNet+CodecNet, C=128, Cy=Cz=64, seed %d); 1 GOP per GPU per step'
                        % ('Low-delay P' if WORKLOAD == 'ldp720' else 'Random Access', GOP_NAME, n_frames, W, H, MODEL['seed']),
            'frames_per_step_per_gpu': n_frames, 'gop': GOP_NAME, 'resolution': '%dx%d' % (W, H),
            'sharding': 'one GOP per rank, no data-path collective' if sharding == 'gop' else
                        'frames of one GOP dealt over the ranks by dependency level, NCCL broadcast of every new 8-bit reconstruction',
            'l2': 'working set per step (activations > 2 GB, 102 MB of frames) exceeds the 126 MB L2'}


# ------------------------------------------------------------------------------ library baseline on the same GPU
def gpu_library_baseline(dev):
    """The reference graph in eager PyTorch (cuDNN / cuBLAS kernels) ON THIS GPU: g_a and g_s of the CodecNet stand-in
    at 1080p through the oracle's torch.nn.functional restatement, fp32 (TF32 convolutions allowed) and bf16 --
    SURVEY.md 2.2's second bar ("the reference GPU path").  Bench-leg use of oracle/."""
    import copy
    import torch
    from aivc_b200 import models
    from oracle import nn_ref as R
    net = models.build_standin(**MODEL).codec_net.codec_net

    def to_dev(m, dtype):
        m = copy.deepcopy(m).to(dev).to(dtype)
        for sub in m.modules():
            if type(sub).__name__ == 'GDN':             # plain tensor attributes, not buffers (misc_layers.py:78-111)
                for a in ('beta_bound', 'gamma_bound', 'pedestal'):
                    setattr(sub, a, getattr(sub, a).to(dev).to(dtype))
        return m

    def time_ms(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    out = {}
    from aivc_b200.codec import latent_dims
    (hy, wy), _ = latent_dims(H, W)
    cases = {'g_a': (net.g_a, (1, 6, H, W)), 'g_s': (net.g_s, (1, net.nb_ft_y + net.out_c_shortcut_y, hy, wy))}
    with torch.no_grad():
        for name, (mod, shape) in cases.items():
            for label, dtype in (('fp32_tf32', torch.float32), ('bf16', torch.bfloat16)):
                m = to_dev(mod, dtype)
                x = torch.rand(shape, device=dev).to(dtype)
                out['%s_%s_ms' % (name, label)] = time_ms(lambda: R.forward_module(m, x))
    return out


# ------------------------------------------------------------------------------ CUDA arm
class Arm:
    """One FrameCodec + the timed legs of the benchmark on it."""

    def __init__(self, precision, net, gop, names, host, resident, out_host, dev, world, rank, local, args):
        import torch
        from aivc_b200 import _lib
        from aivc_b200.codec import FrameCodec
        from aivc_b200.plan import Config
        self.torch, self.L, self._lib = torch, _lib.lib(), _lib
        self.codec = FrameCodec(net, H, W, dev, Config(precision=precision))
        self.precision, self.gop, self.names = precision, gop, names
        self.host, self.resident, self.out_host, self.dev = host, resident, out_host, dev
        self.world, self.rank, self.local, self.args = world, rank, local, args
        self.state = {}

    def step(self, e2e):
        torch, codec, names, state = self.torch, self.codec, self.names, self.state
        if e2e:
            frames = {f: tuple(p.to(self.dev, non_blocking=True) for p in self.host[f]) for f in names}
        else:
            frames = self.resident
        t0 = time.perf_counter()
        bts, rec = codec.encode_gop(frames, self.gop)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        dec = codec.decode_gop(bts, self.gop)
        torch.cuda.synchronize()
        state['enc_s'] = state.get('enc_s', 0.0) + (t1 - t0)
        state['dec_s'] = state.get('dec_s', 0.0) + (time.perf_counter() - t1)
        if e2e:
            for f in names:
                for d, s in zip(self.out_host[f], dec[f]):
                    d.copy_(s, non_blocking=True)
        state['bts'], state['rec'], state['dec'] = bts, rec, dec

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, e2e, profile, steps):
        torch, L, codec = self.torch, self.L, self.codec
        self.barrier()
        l0 = L.aivc_launch_count()
        if profile:
            self._lib.set_profiling(True)
        # per-stage timing needs kernels one at a time: no second stream next to the timed stages
        # (the library likewise drops its two-lane execution while profiling)
        overlap = codec.mof.overlap_shortcut
        codec.mof.overlap_shortcut = codec.codec.overlap_shortcut = overlap and not profile
        codec.lanes_enabled = not profile             # (likewise: one frame in flight while stages are timed)
        sampler = ClockSampler(self.local) if self.rank == 0 else None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            self.step(e2e)
        b.record()
        self.barrier()
        ms = a.elapsed_time(b)
        clocks = sampler.stop() if sampler else None
        prof = None
        if profile:
            out = (C.c_double * 6)()
            self._lib.check(L.aivc_profile_read(out))
            ncls = len(self._lib.KERNEL_CLASSES)
            cls = (C.c_double * (3 * ncls))()
            self._lib.check(L.aivc_profile_read_classes(cls, ncls))
            prof = list(out) + list(cls)
            if self.args.stage_csv and self.rank == 0 and getattr(self, 'dump_csv', True):
                self._lib.check(L.aivc_profile_dump(self.args.stage_csv.encode()))
            self._lib.set_profiling(False)
        codec.mof.overlap_shortcut = codec.codec.overlap_shortcut = overlap
        codec.lanes_enabled = True
        launches = L.aivc_launch_count() - l0
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks, prof, launches

    def check_closed_loop(self, what):
        """decoder output == encoder reconstruction on the data of the last step"""
        for f in self.names:
            for x, y in zip(self.state['rec'][f], self.state['dec'][f]):
                assert self.torch.equal(x, y), 'closed loop broken (%s leg, %s, %s)' % (what, self.precision, f)

    def check_e2e_output(self):
        """the planes that came back to pinned host memory in the e2e leg are the encoder's reconstruction"""
        self.torch.cuda.synchronize()
        for f in self.names:
            for h_, d_ in zip(self.out_host[f], self.state['rec'][f]):
                assert self.torch.equal(h_, d_.cpu()), 'e2e leg returned wrong planes (%s, %s)' % (self.precision, f)


def roofline_of(prof, steps, ms_prof, precision, kernel_classes):
    peak_tf, peak_bw, peak_src = peaks()
    mma_per_flop = 3.0 if precision == 'bf16x3' else 1.0     # tensor-core work per algorithmic FLOP
    peak_alg = peak_tf / mma_per_flop
    tc_ms, tc_fl, tc_n, si_ms, si_fl, si_n = prof[:6]
    ach_all = tc_fl / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    by_kernel, all_ms = {}, sum(prof[6 + 3 * k] for k in range(len(kernel_classes)))
    for k, name in enumerate(kernel_classes):
        k_ms, k_fl, k_n = prof[6 + 3 * k: 9 + 3 * k]
        if k_n:
            tf = k_fl / (k_ms * 1e-3) / 1e12
            by_kernel[name] = {'launches_per_step': k_n / steps, 'avg_launch_us': 1e3 * k_ms / k_n, 'tflops': tf,
                               'frac': tf / peak_alg, 'share_of_stage_time': k_ms / all_ms}
    dom = max((n for n in by_kernel if by_kernel[n]['tflops'] > 0), key=lambda n: by_kernel[n]['share_of_stage_time'])
    ach = by_kernel[dom]['tflops']
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'r02_traffic.json')     # dram bytes per launch from the ncu --set full captures
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(precision, {}).get(dom)
    return {
        'bound': 'tensor', 'achieved': ach, 'peak': peak_alg, 'unit': 'TFLOP/s', 'frac': ach / peak_alg,
        'traffic': traffic,
        # the same numbers in tensor-pipe terms: MMA FLOP/s actually issued against the measured bf16 rate itself
        'mma_per_flop': mma_per_flop, 'tensor_pipe_tflops': mma_per_flop * ach, 'tensor_pipe_peak': peak_tf,
        'peak_source': peak_src + ('; bf16x3 issues three bf16 MMAs (hi.Whi + lo.Whi + hi.Wlo) per algorithmic multiply-add, so the '
                                   'peak for ALGORITHMIC FLOP/s is the measured bf16 rate / 3 (SURVEY.md 8d: "report against '
                                   'the corresponding peak"); tensor-pipe rate = 3 x achieved = %.0f TFLOP/s of %.0f'
                                   % (3 * ach, peak_tf) if precision == 'bf16x3' else ''),
        'kernel': dom + ' (the kernel with the largest share of GPU time)',
        'how': 'algorithmic FLOPs (SURVEY.md 8d) of this kernel\'s launches in K timed steps / sum of their CUDA-event '
               'durations on the launching stream (second pass of the same K steps, an event pair around every '
               'convolution stage); `traffic` = DRAM bytes of one representative launch (ncu --set full, profiles/)',
        'launches_per_step': by_kernel[dom]['launches_per_step'],
        'avg_launch_us': by_kernel[dom]['avg_launch_us'],
        'share_of_stage_time': by_kernel[dom]['share_of_stage_time'],
        'by_kernel': by_kernel,
        'all_tensor_stages': {'tflops': ach_all, 'frac': ach_all / peak_alg, 'stages': int(tc_n)},
        'tc_ms_per_step': tc_ms / steps, 'tc_share_of_step': tc_ms / ms_prof,
        'profiled_ms_per_step': ms_prof / steps,
        'simt_ms_per_step': si_ms / steps, 'simt_tflops': si_fl / max(si_ms, 1e-9) / 1e9,
    }


def run_ours(args):
    import torch
    import torch.distributed as dist
    from aivc_b200 import models, gop as G, _lib

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    if args.sharding == 'frame':
        return run_frame_sharded(args, dev, world, rank, local)

    net = models.build_standin(**MODEL)
    gop = G.generate_gop_struct(GOP_NAME)
    names = sorted(gop, key=lambda f: int(f.split('_')[1]))
    clip = synth_gop(100 + rank, len(names))
    host = {f: tuple(torch.from_numpy(p.reshape(-1)).pin_memory() for p in clip[i]) for i, f in enumerate(names)}
    resident = {f: tuple(p.to(dev) for p in host[f]) for f in names}
    out_host = {f: tuple(torch.empty_like(p).pin_memory() for p in host[f]) for f in names}
    frame_bytes = sum(p.numel() for p in host[names[0]])
    common = (net, gop, names, host, resident, out_host, dev, world, rank, local, args)

    arm = Arm(args.precision, *common)
    for _ in range(args.warmup):
        arm.step(False)
    arm.state['enc_s'] = arm.state['dec_s'] = 0.0
    ms, clocks, _, launches = arm.timed(False, False, args.steps)
    enc_ms, dec_ms = 1e3 * arm.state['enc_s'] / args.steps, 1e3 * arm.state['dec_s'] / args.steps
    arm.check_closed_loop('resident')
    # roofline leg: the same K steps again with a CUDA-event pair around every convolution stage
    # (the event records sit between kernels, which costs a few % of step time, so the headline
    # `value` above is taken without them)
    ms_prof, _, prof, _ = arm.timed(False, True, args.steps)
    total_bytes = sum(len(b) for b in arm.state['bts'].values())
    arm.step(True)
    ms_e2e, _, _, _ = arm.timed(True, False, args.steps)
    arm.check_closed_loop('e2e')
    arm.check_e2e_output()

    frames_per_step = len(names) * world
    value = frames_per_step * args.steps / (ms / 1000.0)
    e2e = frames_per_step * args.steps / (ms_e2e / 1000.0)
    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None,
            'dtype': {'bf16': 'bf16', 'bf16x3': 'bf16x3 (split-bf16 operands on tcgen05, fp32 accumulate: fp32-grade)',
                      'fp32': 'f32'}[args.precision],
            'precision': args.precision,
            'data': 'synthetic', 'config': workload_config(world), 'clocks': clocks,
            'e2e': {'value': e2e, 'unit': 'frames/s',
                    'h2d_bytes_per_step': frame_bytes * len(names), 'd2h_bytes_per_step': frame_bytes * len(names),
                    'output_checked': True,
                    'note': 'frames start in pinned host memory and decoded planes return to pinned host memory '
                            'every step (compared with the encoder\'s reconstruction after the timed region); the '
                            'bitstream (bytes) is produced/consumed on the host in both modes'},
            'gpu_launches': int(launches),
            'roofline': roofline_of(prof, args.steps, ms_prof, args.precision, _lib.KERNEL_CLASSES),
            'bitstream_bytes_per_gop': total_bytes, 'encode_ms_per_gop': enc_ms, 'decode_ms_per_gop': dec_ms,
            'encode_fps': 1e3 * len(names) / enc_ms, 'decode_fps': 1e3 * len(names) / dec_ms,
            'encode_decode_closed_loop': True,
        }
        if world == 1 and not args.quick:
            # parity of THIS engine against the CPU oracle's fixture at the benchmark's resolution (1080p I, P, B;
            # tests/parity_cfg.py, fixture minted by oracle/gen_golden_configs.py)
            from tests import parity_cfg
            keys = ('y_symbols', 'y_mismatches', 'y_index_mismatch_rate', 'y_max_abs_diff', 'z_symbols', 'z_mismatches',
                    'bytes', 'oracle_bytes', 'bytes_delta', 'frames_bytes_identical', 'max_level_diff_subsampled',
                    'max_abs_psnr_delta_db', 'closed_loop_exact')
            case = 'ldp720' if WORKLOAD == 'ldp720' else 'ra1080'
            r = parity_cfg.measure(case, args.precision, dev)
            line['parity'] = dict({k: r[k] for k in keys}, psnr_delta_db=r['max_abs_psnr_delta_db'],
                                  vs='CPU oracle fixture tests/golden/cfg_%s.npz (%s; stand-in with live hyperprior)'
                                  % (case, '720p I, P, P' if case == 'ldp720' else '1080p I, P, B'))
            # the other tensor-core precision next to the headline one, same workload, same run
            other = 'bf16' if args.precision != 'bf16' else 'bf16x3'
            del arm
            torch.cuda.empty_cache()
            alt = Arm(other, *common)
            alt.dump_csv = False                      # (--stage-csv holds the headline arm's stages)
            alt.step(False)
            alt.step(False)
            ms_a, _, _, _ = alt.timed(False, False, args.steps)
            alt.check_closed_loop('resident')
            alt.step(True)
            ms_ae, _, _, _ = alt.timed(True, False, args.steps)
            alt.check_e2e_output()
            ms_ap, _, prof_a, _ = alt.timed(False, True, args.steps)
            rf = roofline_of(prof_a, args.steps, ms_ap, other, _lib.KERNEL_CLASSES)
            ra = parity_cfg.measure(case, other, dev)
            line['other_precision'] = {
                'precision': other, 'value': frames_per_step * args.steps / (ms_a / 1000.0),
                'e2e': frames_per_step * args.steps / (ms_ae / 1000.0), 'unit': 'frames/s',
                'roofline': {k: rf[k] for k in ('kernel', 'achieved', 'peak', 'frac', 'all_tensor_stages')},
                'parity': {k: ra[k] for k in keys},
                'note': 'bf16: plain bf16 operands -- the fastest mode, ~1 % of the latent indices differ from fp32 '
                        'arithmetic; bf16x3: the fp32-faithful default'}
            del alt
            torch.cuda.empty_cache()
            line['gpu_library_baseline'] = dict(gpu_library_baseline(dev), note='eager PyTorch (cuDNN) on this GPU: g_a / g_s of '
                                                'the CodecNet stand-in at 1080p, ms per call; the fused plans of this repo take '
                                                'the times in roofline.by_kernel')
        if world == 1 and not args.no_cpu_baseline:
            fps, threads, sample, _ = cpu_reference_sample(25.0)
            line['cpu_baseline'] = {'value': fps, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
                                    'sample': sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_frame_sharded(args, dev, world, rank, local):
    """--sharding frame: ONE GOP per step, its frames dealt over the ranks by dependency level, every new 8-bit
    reconstruction broadcast over NCCL before the next level (BASELINE.json configs[3]).  Latency
    mode: value = frames of the GOP / time of the slowest rank."""
    import torch
    import torch.distributed as dist
    from aivc_b200 import models, gop as G, sharding, _lib
    from aivc_b200.codec import FrameCodec
    from aivc_b200.plan import Config
    net = models.build_standin(**MODEL)
    gop = G.generate_gop_struct(GOP_NAME)
    names = sorted(gop, key=lambda f: int(f.split('_')[1]))
    codec = FrameCodec(net, H, W, dev, Config(precision=args.precision))
    clip = synth_gop(100, len(names))                    # every rank holds the whole GOP's source frames
    frames = {f: tuple(torch.from_numpy(p.reshape(-1)).to(dev) for p in clip[i]) for i, f in enumerate(names)}
    L = _lib.lib()
    stats = {}

    def step():
        bts, rec = sharding.encode_gop_frame_parallel(codec, frames, gop, rank, world, stats=stats)
        dec = sharding.decode_gop_frame_parallel(codec, bts, gop, rank, world, stats=stats)
        return bts, rec, dec

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    stats.clear()
    l0 = L.aivc_launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        bts, rec, dec = step()
    b.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    for f in names:
        for x, y in zip(rec[f], dec[f]):
            assert torch.equal(x, y), 'closed loop broken on ' + f
    import hashlib
    md5 = hashlib.md5(b''.join(bts[f] for f in names)).hexdigest()
    if rank == 0:
        value = len(names) * args.steps / (ms / 1000.0)
        print(json.dumps({
            'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': args.precision, 'precision': args.precision, 'data': 'synthetic',
            'config': workload_config(world, 'frame'), 'clocks': clocks,
            'gpu_launches': int(L.aivc_launch_count() - l0),
            'gop_latency_ms': ms / args.steps,
            'broadcast_ms_per_step': 1e3 * stats.get('bcast_s', 0.0) / args.steps,
            'broadcasts_per_step': stats.get('bcasts', 0) / args.steps,
            'bitstream_md5': md5, 'bitstream_bytes_per_gop': sum(len(bts[f]) for f in names),
            'encode_decode_closed_loop': True,
            'note': 'one GOP per step over all ranks (strong scaling, latency mode); the bitstream md5 is identical for every N'}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default='bf16x3', choices=['bf16x3', 'bf16', 'fp32'],
                    help='bf16x3 (default): fp32-faithful split-bf16 tensor-core mode; bf16: fastest; fp32: exact SIMT engine')
    ap.add_argument('--sharding', default='gop', choices=['gop', 'frame'],
                    help='gop: one GOP per rank (throughput, no collective); frame: one GOP over all ranks (latency, NCCL broadcasts)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--quick', action='store_true', help='skip the parity / other-precision / library-baseline legs')
    ap.add_argument('--stage-csv', default='', help='dump per-stage CUDA-event timings of the timed region')
    ap.add_argument('--workload', default='ra1080', choices=['ra1080', 'ldp720'],
                    help='ra1080: BASELINE configs[2] (default, the headline); ldp720: configs[1] (1280x720 low-delay P, GOP 8)')
    ap.add_argument('--model-seed', type=int, default=None,
                    help='stand-in weights seed (configs[4]: seeds 1..7 stand for the models ms_ssim-1..7)')
    args = ap.parse_args()
    set_workload(args.workload, args.model_seed)
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
