#!/bin/bash
# one ncu --set full capture of k_model (one band launch in the middle of a frame batch); env PROBE_* select the content
TAG=${1:-x}
B=${2:-64}
KIND=${3:-grain}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_model -s 5 -c 1 -o gpurun_out/${TAG}_k_model -f \
    python tools/probe_content.py $B $KIND > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/${TAG}_*
