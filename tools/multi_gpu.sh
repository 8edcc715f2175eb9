#!/bin/bash
# usage: tools/multi_gpu.sh N   -- weak-scaling (one GOP per rank) and frame-sharded (one GOP over all ranks) bench on N GPUs
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "1" ]; then TR="python"; fi
$TR bench.py --gpus $N --steps 2 --warmup 2 --quick --no-cpu-baseline > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/mg_${N}.err || tail -5 gpurun_out/mg_${N}.err
$TR bench.py --gpus $N --steps 2 --warmup 1 --sharding frame > gpurun_out/r02_frame_sharding_${N}gpu.json 2> gpurun_out/mgf_${N}.err || tail -5 gpurun_out/mgf_${N}.err
python - <<PY
import json
for f in ("gpurun_out/r02_bench_${N}gpu.json", "gpurun_out/r02_frame_sharding_${N}gpu.json"):
    try:
        b = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "value %.2f fps" % b["value"], "ms/step %.1f" % b["ms_per_step"], {k: b.get(k) for k in ("gop_latency_ms", "broadcast_ms_per_step", "broadcasts_per_step", "bitstream_md5")})
    except Exception as e:
        print(f, "FAILED", e)
PY
