#!/bin/bash
# Round-2 evidence on one B200 box -> gpurun_out/ (copy what is to be judged into profiles/).
#   tools/profile_round.sh [part ...]     parts: tests bench ref ldp sweep parity fp32 ncu   (default: all)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
parts="${@:-tests bench ref ldp sweep parity fp32 ncu}"
last() { grep '^{' "$1" | tail -1; }
for part in $parts; do
case $part in
tests) timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_gpu_tests.txt; cat gpurun_out/r02_gpu_tests.txt ;;
bench) timeout 900 python bench.py --steps 3 --warmup 3 --stage-csv gpurun_out/r02_stage_times.csv > gpurun_out/r02_bench.raw 2> gpurun_out/r02_bench.err; last gpurun_out/r02_bench.raw > gpurun_out/r02_bench.json; tail -2 gpurun_out/r02_bench.err ;;
ref) timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_ref.raw 2> gpurun_out/r02_ref.err; last gpurun_out/r02_ref.raw > gpurun_out/r02_bench_reference_arm.json ;;
ldp) timeout 600 python bench.py --workload ldp720 --steps 3 --warmup 3 > gpurun_out/r02_ldp.raw 2> gpurun_out/r02_ldp.err; last gpurun_out/r02_ldp.raw > gpurun_out/r02_bench_ldp720.json ;;
sweep) : > gpurun_out/r02_bench_sweep.jsonl; for seed in 1 4 7; do timeout 300 python bench.py --model-seed $seed --steps 2 --warmup 2 --quick --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1 >> gpurun_out/r02_bench_sweep.jsonl; done ;;
parity) timeout 600 python tools/parity_configs.py gpurun_out/r02_parity_configs.json 2>&1 | tail -9 ;;
fp32) timeout 600 python tools/fp32_engine_fps.py 2>/dev/null | tail -1 > gpurun_out/r02_fp32_engine.json; cat gpurun_out/r02_fp32_engine.json ;;
ncu)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_ncu_launches_3frames.csv python tools/frame_once.py 1 bf16x3 > gpurun_out/ncu_list.log 2>&1
  for spec in "conv3x3_tc_kernel:c3_128_270:r02_conv3x3_x3" "conv3x3_tc_gdn_kernel:c3igdn_res_544:r02_conv3x3_gdn_x3" "tconv3x3_tc_kernel:up3_128_270:r02_tconv3x3_x3"; do
    IFS=: read k case out <<< "$spec"
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/$out python tools/bench_layer.py --precision bf16x3 --cases $case --iters 1 --reps 2 > gpurun_out/ncu_$out.log 2>&1
    ncu -i gpurun_out/$out.ncu-rep --page raw --csv > gpurun_out/${out}_ncu_raw.csv 2>/dev/null
  done ;;
esac
done
ls -la gpurun_out | tail -30
