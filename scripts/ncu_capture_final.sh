#!/bin/bash
# Final-binary refresh of the launch list and the decoder-step full capture (see ncu_capture.sh for the full set).
OUT=gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 140 --csv --log-file $OUT/launches_final.csv $B > $OUT/launches_final.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decoder_step2 -s 100 -c 1 -o $OUT/full_decoder_step2_final -f $B > $OUT/full_decoder_final.log 2>&1
ls -la $OUT | tail -5
