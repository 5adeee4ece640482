#!/bin/bash
# Phased fused pass Y+X (PAWB200_FFT_FUSED=2): correctness on the FFT/parity tests, then A/B against the stand-alone passes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_c.log 2>&1
tail -3 gpurun_out/pytest_c.log
PAWB200_FFT_FUSED=2 timeout 600 python -m pytest tests/test_gpu_fft_paths.py tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q > gpurun_out/pytest_fused2.log 2>&1
tail -3 gpurun_out/pytest_fused2.log
b2() { timeout 300 python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu 2> gpurun_out/$1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms_per_step'].items() if v>0}, d['checksum'])"; }
b3() { timeout 400 python bench.py --nband 512 --steps 3 --warmup 3 --no-cpu --no-secondary 2> gpurun_out/$1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms_per_step'].items() if v>0}, d['checksum'])"; }
PAWB200_FFT_PF=0 b2 cfg2_pf0
PAWB200_FFT_PF=1 b2 cfg2_pf1
PAWB200_FFT_FUSED=2 b2 cfg2_fused2_60
PAWB200_FFT_FUSED=2 PAWB200_FFT_RING_BYTES=42000000 b2 cfg2_fused2_40
PAWB200_FFT_FUSED=2 PAWB200_FFT_RING_BYTES=94000000 b2 cfg2_fused2_90
PAWB200_FFT_FUSED=2 PAWB200_FFT_PF=0 b2 cfg2_fused2_60_pf0
PAWB200_FFT_PF=0 b3 cfg3_pf0
PAWB200_FFT_PF=1 b3 cfg3_pf1
PAWB200_FFT_FUSED=2 b3 cfg3_fused2_60
PAWB200_FFT_FUSED=2 PAWB200_FFT_RING_BYTES=94000000 b3 cfg3_fused2_90
PAWB200_FFT_FUSED=2 PAWB200_FFT_RING_BYTES=40000000 b3 cfg3_fused2_40
PAWB200_FFT_FUSED=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fft_pass" -s 6 -c 2 -o gpurun_out/r02c_fused2_cfg2 python bench.py --config cfg2 --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_f2.log 2>&1
tail -1 gpurun_out/ncu_f2.log | cut -c1-150
bash scripts/ncu_export.sh gpurun_out/r02c_fused2_cfg2
B="python bench.py --steps 1 --warmup 0 --no-cpu --no-secondary --nband 512"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sphere_project_real -s 2 -c 3 -o gpurun_out/r02c_project_cfg3 $B > gpurun_out/ncu_proj.log 2>&1
bash scripts/ncu_export.sh gpurun_out/r02c_project_cfg3
du -sh gpurun_out
