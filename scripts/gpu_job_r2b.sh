#!/bin/bash
# A/B of the L2 prefetch in the pruned FFT passes + the MINB=4 build of the radix<=10 class; source-level captures.
mkdir -p gpurun_out
run() { # name, env..., -- args
  name=$1; shift
  env "$@" > /dev/null 2>&1
}
b2() { python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu 2> gpurun_out/$1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms_per_step'].items() if v>0})"; }
b3() { python bench.py --nband 512 --steps 3 --warmup 3 --no-cpu --no-secondary 2> gpurun_out/$1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms_per_step'].items() if v>0})"; }
PAWB200_FFT_PF=0 b2 cfg2_pf0
PAWB200_FFT_PF=1 b2 cfg2_pf1
PAWB200_FFT_PF=0 b3 cfg3_pf0
PAWB200_FFT_PF=1 b3 cfg3_pf1
cp pawpyseed_b200/libpawb200.so /tmp/libpawb200_base.so
cp variants/libpawb200_m4.so pawpyseed_b200/libpawb200.so
PAWB200_FFT_PF=0 b2 cfg2_m4_pf0
PAWB200_FFT_PF=1 b2 cfg2_m4_pf1
PAWB200_FFT_PF=1 b3 cfg3_m4_pf1
cp /tmp/libpawb200_base.so pawpyseed_b200/libpawb200.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fft_pass" -s 6 -c 3 -o gpurun_out/r02b_fft_cfg2 python bench.py --config cfg2 --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_fft2.log 2>&1
tail -1 gpurun_out/ncu_fft2.log | cut -c1-150
bash scripts/ncu_export.sh gpurun_out/r02b_fft_cfg2
B="python bench.py --steps 1 --warmup 0 --no-cpu --no-secondary --nband 512"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fft_pass" -s 3 -c 3 -o gpurun_out/r02b_fft_cfg3 $B > gpurun_out/ncu_fft3.log 2>&1
tail -1 gpurun_out/ncu_fft3.log | cut -c1-150
bash scripts/ncu_export.sh gpurun_out/r02b_fft_cfg3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sphere_project_real -s 2 -c 3 -o gpurun_out/r02b_project_cfg3 $B > gpurun_out/ncu_proj.log 2>&1
tail -1 gpurun_out/ncu_proj.log | cut -c1-150
bash scripts/ncu_export.sh gpurun_out/r02b_project_cfg3
du -sh gpurun_out
