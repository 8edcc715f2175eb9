#!/bin/bash
# A/B of the working-tree library against aivc_b200/libaivc_b200_base.so on the same box
CASES=${1:-c3_128_540,c3_128_270,res_128_270}
for i in 1 2; do
echo "== base";  AIVC_B200_LIB=$PWD/aivc_b200/libaivc_b200_base.so timeout 100 python tools/bench_layer.py --cases $CASES 2>&1 | grep -v "^$"
echo "== new";  timeout 100 python tools/bench_layer.py --cases $CASES 2>&1 | grep -v "^$"
done
