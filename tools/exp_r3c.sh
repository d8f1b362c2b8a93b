#!/bin/bash
# k_model CTA size with the serial one-lane-per-slot S2 (16 / 24 / 32 class warps)
mkdir -p gpurun_out
L=gpurun_out/r3c.log
: > $L
run() { echo "== B=${B:-128} ${K:-grain} $*" >> $L; env "$@" python tools/probe_content.py ${B:-128} ${K:-grain} 2>&1 | grep "^B=\|kernel" | tail -2 >> $L; }
for T in 768 1024 512; do
  echo "#### threads $T" >> $L
  B200_MODEL_THREADS=$T python __graft_entry__.py -f > /dev/null 2>&1
  (timeout 900 python -m pytest tests/test_ffv1_gpu.py -m gpu -x -q 2>&1 | tail -1) >> $L
  PROBE_KERNELS=1 run X=1
  K=flat PROBE_KERNELS=1 run X=1
  PROBE_W=2048 PROBE_H=1556 PROBE_LAYOUT=2 PROBE_SLICES=4 K=grain PROBE_KERNELS=1 run X=1
done
cat $L
