#!/bin/bash
# Leaner line bodies (32-bit offsets, incremental line indices, strided phase-2 stores): tests, A/B numbers, captures, e2e timeline.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_d.log 2>&1
tail -3 gpurun_out/pytest_d.log
b2() { timeout 300 python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu 2> gpurun_out/$1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms_per_step'].items() if v>0}, d['checksum'])"; }
b3() { timeout 400 python bench.py --nband 512 --steps 3 --warmup 3 --no-cpu --no-secondary 2> gpurun_out/$1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms_per_step'].items() if v>0}, d['checksum'])"; }
PAWB200_FFT_PF=1 b2 cfg2_pf1
PAWB200_FFT_PF=0 b2 cfg2_pf0
PAWB200_FFT_PF=1 b3 cfg3_pf1
PAWB200_FFT_PF=0 b3 cfg3_pf0
B="python bench.py --steps 1 --warmup 0 --no-cpu --no-secondary --nband 512"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fft_pass" -s 3 -c 3 -o gpurun_out/r02d_fft_cfg3 $B > gpurun_out/ncu_fft3.log 2>&1
bash scripts/ncu_export.sh gpurun_out/r02d_fft_cfg3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fft_pass" -s 6 -c 3 -o gpurun_out/r02d_fft_cfg2 python bench.py --config cfg2 --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_fft2.log 2>&1
bash scripts/ncu_export.sh gpurun_out/r02d_fft_cfg2
# one rank of the 8-GPU job: 1 block, 2 host threads, device trace
PAWB200_TRACE=1 timeout 600 python scripts/e2e_timeline.py 3 cfg3 1 2 > gpurun_out/timeline_cfg3_1block.log 2> gpurun_out/timeline_cfg3_1block.trace
tail -4 gpurun_out/timeline_cfg3_1block.log
du -sh gpurun_out
