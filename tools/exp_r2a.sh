#!/bin/bash
# round-2 experiment A: where does k_model's time go (phase cycles), chain threshold 1, baseline numbers
mkdir -p gpurun_out
L=gpurun_out/r2a.log
: > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $L
echo "== baseline B=64 grain" >> $L
python tools/probe_content.py 64 grain >> $L 2>&1
echo "== serial timing B=64" >> $L
PROBE_KERNELS=1 python tools/probe_content.py 64 grain >> $L 2>&1
echo "== phase timing" >> $L
B200_PHASE_TIMING_BUILD=1 python __graft_entry__.py -f >> $L 2>&1
B200_PHASE_TIMING=1 B200_SERIAL=1 python tools/probe_content.py 16 grain >> $L 2>&1
for cm in 1 2; do
  echo "== CHAIN_MIN=$cm" >> $L
  B200_EXTRA_NVCC="-DB200_CHAIN_MIN=$cm" python __graft_entry__.py -f >> $L 2>&1
  PROBE_KERNELS=1 python tools/probe_content.py 64 grain >> $L 2>&1
done
tail -60 $L
