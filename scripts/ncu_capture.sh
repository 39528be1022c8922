#!/bin/bash
# Round-1 profile captures (run on the GPU box under gpurun; outputs land in gpurun_out/):
#   launch list of one full step, then `--set full` captures of the three hand-written hot kernels.
set -x
OUT=gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 140 --csv --log-file $OUT/launches.csv $B > $OUT/launches.log 2>&1
for k in decoder_step2 attention_tc gemm_tc; do
  skip=20; [ $k = decoder_step2 ] && skip=100
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o $OUT/full_$k -f $B > $OUT/full_$k.log 2>&1
done
ls -la $OUT
