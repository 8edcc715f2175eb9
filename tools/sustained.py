"""Back-to-back launches of one stage for a few seconds, with nvidia-smi clock/power sampling."""
import os, subprocess, sys, tempfile, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aivc_b200.layers as M
from aivc_b200.plan import Plan, Config
from tools.bench_layer import CASES

name = sys.argv[1] if len(sys.argv) > 1 else 'c3_128_540'
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
mk, cin, h, w = CASES[name]
dev = torch.device('cuda:0')
plan = Plan(mk().eval(), h, w, cin, dev, Config(precision='bf16'))
plan.src.buf.t.normal_()
for _ in range(20):
    plan.run()
torch.cuda.synchronize()
f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
p = subprocess.Popen(['nvidia-smi', '-i', '0', '--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown',
                      '--format=csv,noheader,nounits', '-lms', '100'], stdout=f)
n = 0
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
a.record()
while time.time() - t0 < secs:
    for _ in range(200):
        plan.run()
    n += 200
    torch.cuda.synchronize()
b.record()
torch.cuda.synchronize()
p.terminate(); p.wait()
ms = a.elapsed_time(b)
f.seek(0)
rows = [l.strip().split(', ') for l in f.read().splitlines() if l.strip()]
clk = sorted(float(r[0]) for r in rows[3:]) or [0]
pw = sorted(float(r[1]) for r in rows[3:]) or [0]
cap = sum(1 for r in rows[3:] if r[2].startswith('Active'))
print('%s: %d launches, %.1f us each, %.1f TFLOP/s sustained; sm clock median %.0f MHz (min %.0f), power median %.0f W (max %.0f), sw_power_cap active in %d/%d samples'
      % (name, n, ms * 1e3 / n, plan.flops() * n / (ms * 1e-3) / 1e12, clk[len(clk) // 2], clk[0], pw[len(pw) // 2], pw[-1], cap, len(rows) - 3))
