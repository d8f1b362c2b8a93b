#!/bin/bash
# round-2 experiment H: FLAC LPC on the GPU, full GPU suite, CLI throughput with mmap writes + parallel pinning
mkdir -p gpurun_out
L=gpurun_out/r2h.log
: > $L
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) >> $L
echo "== cli throughput 256 files, 64 in flight" >> $L
timeout 600 python tools/cli_throughput.py 256 >> $L 2>&1
echo "== cli throughput 512 files, 64 in flight" >> $L
timeout 600 python tools/cli_throughput.py 512 2>&1 | grep "^run" >> $L
timeout 600 python bench.py --config 4 --no-cpu > gpurun_out/r2h_bench4.json 2>> $L
cat $L
cut -c1-600 gpurun_out/r2h_bench4.json
