set -x
cd /root/repo
timeout 300 python bench.py --steps 3 --warmup 3 --stage-csv gpurun_out/r01_stage_times.csv > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
timeout 200 python tools/parity_report.py > gpurun_out/r01_parity_report.json 2> gpurun_out/parity.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r01_ncu_launches_3frames.csv python tools/frame_once.py 1 > gpurun_out/ncu_list.log 2>&1
