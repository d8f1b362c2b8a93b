#!/bin/bash
# Runs on the GPU box (gpurun): GPU tests, the bench line, the reference arm, the ncu launch list of the bench command
# and one `ncu --set full` capture per kernel. Outputs under gpurun_out/<tag>_*; summarise with profiles/summarize.py <tag>.
TAG=${1:-r1}
F=${2:-128}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/${TAG}_gpu_tests.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 400 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 420 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --frames $F --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_launches.log 2>&1
for K in k_model k_range k_emit k_pack; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -o gpurun_out/${TAG}_$K -f \
      python bench.py --frames $F --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_$K.log 2>&1
done
tail -3 gpurun_out/${TAG}_gpu_tests.log
cut -c1-400 gpurun_out/${TAG}_bench.json
cut -c1-300 gpurun_out/${TAG}_bench_reference.json
ls -la gpurun_out | tail -12
