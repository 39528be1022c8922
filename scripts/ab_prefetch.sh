set -x
B="python bench.py --headline-only --no-cpu-baseline"
for pf in 0 63 1 2 4; do MOONSHINE_B200_PREFETCH=$pf $B > gpurun_out/r2e_tiny32_pf$pf.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r2e_tiny32_pf$pf.json').read().splitlines()[-1]);print('tiny32 pf=$pf',round(d['value'],1),d['stage_ms']['decode_launch_us'],round(d['e2e']['value'],1))"; done
for pf in 0 63; do MOONSHINE_B200_PREFETCH=$pf $B --batch 1 > gpurun_out/r2e_tiny1_pf$pf.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r2e_tiny1_pf$pf.json').read().splitlines()[-1]);print('tiny1 pf=$pf',round(d['value'],1),d['stage_ms']['decode_launch_us'])"; done
for pf in 0 63 8 16 48; do MOONSHINE_B200_PREFETCH=$pf $B --model base --batch 256 --steps 3 --warmup 2 > gpurun_out/r2e_base256_pf$pf.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r2e_base256_pf$pf.json').read().splitlines()[-1]);print('base256 pf=$pf',round(d['value'],1),d['stage_ms']['decode_launch_us'])"; done
for pf in 0 63; do MOONSHINE_B200_PREFETCH=$pf $B --model base_streaming --batch 64 --steps 3 --warmup 2 > gpurun_out/r2e_bs64_pf$pf.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r2e_bs64_pf$pf.json').read().splitlines()[-1]);print('bs64 pf=$pf',round(d['value'],1),d['stage_ms']['decode_launch_us'])"; done
MOONSHINE_B200_PROF=30 python scripts/prof_step.py tiny 32 2> gpurun_out/r2e_v4_prof_raw.txt >/dev/null; python scripts/prof_v3.py gpurun_out/r2e_v4_prof_raw.txt --v4 2>/dev/null | head -8 | cut -c1-900
timeout 600 python scripts/ncu_traffic.py > gpurun_out/r2e_ncu_traffic.log 2>&1; cp profiles/r2_decoder_traffic* gpurun_out/; cat gpurun_out/r2e_ncu_traffic.log | cut -c1-300
