#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r2t}
L=gpurun_out/$TAG.log
: > $L
(timeout 900 python -m pytest tests/test_ffv1_gpu.py -m gpu -x -q 2>&1 | tail -2) >> $L
run() { echo "== B=${B:-128} ${K:-grain} $*" >> $L; env "$@" python tools/probe_content.py ${B:-128} ${K:-grain} 2>&1 | grep "^B=\|kernel" | tail -2 >> $L; }
PROBE_KERNELS=1 run X=1
K=flat PROBE_KERNELS=1 run X=1
PROBE_W=2048 PROBE_H=1556 PROBE_LAYOUT=2 PROBE_SLICES=4 K=grain PROBE_KERNELS=1 run X=1
K=zero B=32 PROBE_KERNELS=1 run X=1
K=white B=32 PROBE_KERNELS=1 run X=1
cat $L
bash tools/exp_ncu_model.sh $TAG 64 grain
