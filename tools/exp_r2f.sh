#!/bin/bash
mkdir -p gpurun_out
B200_SERIAL=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_range -s 6 -c 1 -o gpurun_out/r2f_k_range -f python tools/probe_content.py 128 grain > gpurun_out/r2f_ncu.log 2>&1
tail -5 gpurun_out/r2f_ncu.log
ls -la gpurun_out/r2f*
