#!/bin/bash
# 3-stage projection kernel (3 CTAs/SM) + leaner forward passes: tests and A/B numbers.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_e.log 2>&1
tail -3 gpurun_out/pytest_e.log
b2() { timeout 300 python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu 2> gpurun_out/$1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms_per_step'].items() if v>0}, d['checksum'])"; }
b3() { timeout 400 python bench.py --nband 512 --steps 3 --warmup 3 --no-cpu --no-secondary 2> gpurun_out/$1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms_per_step'].items() if v>0}, d['checksum'])"; }
b2 cfg2
b3 cfg3
B="python bench.py --steps 1 --warmup 0 --no-cpu --no-secondary --nband 512"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sphere_project_real -s 2 -c 3 -o gpurun_out/r02e_project_cfg3 $B > gpurun_out/ncu_proj.log 2>&1
bash scripts/ncu_export.sh gpurun_out/r02e_project_cfg3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zgemm_abh_kernel -s 1 -c 1 -o gpurun_out/r02e_zgemm_cfg3 $B > gpurun_out/ncu_zgemm.log 2>&1
bash scripts/ncu_export.sh gpurun_out/r02e_zgemm_cfg3
du -sh gpurun_out
