#!/bin/bash
# ncu --set full of k_decode on a smaller geometry (1920x1080 16-bit, 24 slices, B frames): source-level stall samples
TAG=${1:-dec2}
B=${2:-16}
SPW=${3:-1}
mkdir -p gpurun_out
PROBE_W=1920 PROBE_H=1080 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decode -s 1 -c 1 -o gpurun_out/${TAG}_k_decode -f \
    python tools/probe_decode.py $B grain $SPW > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log



ls -la gpurun_out/${TAG}_*
