#!/bin/bash
# Round evidence on one B200 box: GPU tests, bench (+CPU baseline), reference arm, ncu launch list of a
# 3-frame encode+decode, and ncu --set full captures of the persistent kernels.  Outputs -> gpurun_out/.
set -x
cd "$(dirname "$0")/.."
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r01_gpu_tests.txt
timeout 400 python bench.py --steps 3 --warmup 3 --stage-csv gpurun_out/r01_stage_times.csv > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference_arm.json 2> gpurun_out/r01_ref.err
timeout 200 python tools/parity_report.py > gpurun_out/r01_parity_report.json 2> gpurun_out/parity.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r01_ncu_launches_3frames.csv python tools/frame_once.py 1 > gpurun_out/ncu_list.log 2>&1
for spec in "conv3x3_tc_kernel:c3_128_270:r01_conv3x3" "conv3x3_tc_gdn_kernel:c3igdn_res_544:r01_conv3x3_gdn" "conv1x1_tc_kernel:att_128_270:r01_conv1x1" "tconv3x3_tc_kernel:up3_128_270:r01_tconv3x3"; do
  IFS=: read k case out <<< "$spec"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/$out python tools/bench_layer.py --cases $case --iters 1 --reps 2 > gpurun_out/ncu_$out.log 2>&1
  ncu -i gpurun_out/$out.ncu-rep --page raw --csv > gpurun_out/${out}_ncu_raw.csv 2>/dev/null
done
ls -la gpurun_out | tail -20
