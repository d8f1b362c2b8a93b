#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3) > gpurun_out/r4f.log
(timeout 900 python tools/cli_throughput.py 256 all 2>&1 | grep -v "^b200enc:" | tail -12) >> gpurun_out/r4f.log
cat gpurun_out/r4f.log
