#!/bin/bash
# Runs on the GPU box (gpurun): GPU tests, the bench line of every BASELINE config, the reference arm, the ncu launch list
# of the bench command and one `ncu --set full` capture per kernel (k_flac from the config-4 command).
# Outputs under gpurun_out/<tag>_*; summarise with profiles/summarize.py <tag>.
TAG=${1:-r2}
F=${2:-128}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/${TAG}_gpu_tests.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 400 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
for C in 2 4 5; do
  timeout 400 python bench.py --config $C --cpu-seconds 6 > gpurun_out/${TAG}_bench_config$C.json 2>> gpurun_out/${TAG}_bench.err
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 420 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --frames $F --steps 2 --warmup 1 --no-cpu --no-check --no-decode > gpurun_out/${TAG}_ncu_launches.log 2>&1
for K in k_model k_range k_emit k_pack; do
  SK=40; [ $K = k_pack ] && SK=1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s $SK -c 1 -o gpurun_out/${TAG}_$K -f \
      python bench.py --frames $F --steps 1 --warmup 1 --no-cpu --no-check --no-decode > gpurun_out/${TAG}_ncu_$K.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_flac -s 2 -c 1 -o gpurun_out/${TAG}_k_flac -f \
    python bench.py --config 4 --steps 1 --warmup 1 --no-cpu --no-check --no-decode > gpurun_out/${TAG}_ncu_k_flac.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_decode -c 1 -o gpurun_out/${TAG}_k_decode -f \
    python bench.py --frames 32 --steps 1 --warmup 1 --no-cpu --no-check > gpurun_out/${TAG}_ncu_k_decode.log 2>&1
for K in k_md5 k_padding; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/${TAG}_$K -f \
      python bench.py --config 2 --frames 64 --steps 1 --warmup 1 --no-cpu --no-check > gpurun_out/${TAG}_ncu_$K.log 2>&1
done
tail -3 gpurun_out/${TAG}_gpu_tests.log
for f in gpurun_out/${TAG}_bench*.json; do echo $f; cut -c1-300 $f; done
ls -la gpurun_out/${TAG}_*
