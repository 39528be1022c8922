#!/bin/bash
# Round-2 Nsight Compute evidence (run on the GPU box; numbers under ncu are evidence, never bench values):
#   1. launch lists (gpu__time_duration per launch, clocks untouched) of one transcription per operating point
#   2. one --set full capture of each kernel on the path
# Outputs land in gpurun_out/ncu_r2/ ; the summaries are made from them by scripts/ncu_summarise.py
set -u
export MOONSHINE_B200_V4_COOP=0     # ncu cannot replay a clustered + cooperative launch
O=gpurun_out/ncu_r2; mkdir -p $O
for cfg in "tiny 32" "base 256" "base_streaming 64"; do set -- $cfg
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_$1_b$2.csv \
    python scripts/prof_step.py $1 $2 > /dev/null 2>&1
done
cap() {  # name regex model batch [skip]
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s ${5:-0} -c 1 -f -o $O/$1 \
    python scripts/prof_step.py $3 $4 > /dev/null 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.raw.csv 2>/dev/null
}
cap decoder4_tiny_b32 decoder_step4_kernel tiny 32 30
cap decoder3_base_b256 decoder_step3_kernel base 256 30
cap gemm_tc_bs64 gemm_tc_kernel base_streaming 64 1
cap gemm_planes_fc1_tiny_b32 gemm_planes_kernel tiny 32 4
cap gemm_planes_persistent_fc1_base_b256 gemm_planes_persistent_kernel base 256 4
cap layernorm_planes_tiny_b32 layernorm_planes_kernel tiny 32 2
cap groupnorm_im2col_tiny_b32 groupnorm_im2col_planes_kernel tiny 32 0
cap attention_tc_tiny_b32 attention_tc_kernel tiny 32 2
cap conv1_tiny_b32 conv1_tanh_kernel tiny 32 0
cap layernorm_tiny_b32 layernorm_kernel tiny 32 0
cap stream_frames_bs64 stream_frames_kernel base_streaming 64 0
cap attention_tc_bs64 attention_tc_kernel base_streaming 64 2
rm -f $O/*.ncu-rep.tmp; rm -f $O/decoder3_base_b256.ncu-rep  # 12 MB: the raw page is kept
ls -la $O
