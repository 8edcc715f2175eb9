#!/bin/bash
# A/B of the working-tree library against aivc_b200/libaivc_b200_base.so on the same box (box-to-box variance is
# ~5 %, more than most single kernel changes).  Build the base from the last commit first:
#   git stash; make -C aivc_b200/csrc; cp aivc_b200/libaivc_b200.so aivc_b200/libaivc_b200_base.so; git stash pop; make -C aivc_b200/csrc
# (same C ABI required: aivc_conv_op must not have changed in between)
CASES=${1:-c3_128_540,c3_128_270,res_128_270}
for i in 1 2; do
echo "== base";  AIVC_B200_LIB=$PWD/aivc_b200/libaivc_b200_base.so timeout 100 python tools/bench_layer.py --cases $CASES 2>&1 | grep -v "^$"
echo "== new";  timeout 100 python tools/bench_layer.py --cases $CASES 2>&1 | grep -v "^$"
done
