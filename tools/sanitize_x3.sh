#!/bin/bash
# compute-sanitizer passes over the split-bf16 kernels (memcheck: whole x3 engine suite; racecheck / synccheck: a subset)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_engine_x3.py tests/test_gpu_engine_tc.py -q -x -m gpu > gpurun_out/r02_san_x3_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_san_x3_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_engine_x3.py -q -x -m gpu -k "(size1 and (cheng_down_128 or up3_no_128 or conv3_s1_gdn_128 or cheng_plain_128)) or (size3 and cheng_plain_128)" > gpurun_out/r02_san_x3_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_san_x3_racecheck.txt
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_engine_x3.py -q -x -m gpu -k "size1 and (cheng_down_128 or up3_no_128 or conv3_s1_gdn_128 or attention_64)" > gpurun_out/r02_san_x3_synccheck.txt 2>&1; echo "synccheck rc=$?" >> gpurun_out/r02_san_x3_synccheck.txt
for f in memcheck racecheck synccheck; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" gpurun_out/r02_san_x3_$f.txt | tail -4; done
