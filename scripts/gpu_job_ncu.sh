#!/bin/bash
# ncu evidence for profiles/: launch list of one default bench step + full-set captures of the dominant kernels.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-cpu --no-secondary"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_cfg3.csv $B > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zgemm_abh_kernel -s 1 -c 2 -o gpurun_out/r02_zgemm_cfg3 $B > gpurun_out/ncu_zgemm.log 2>&1
tail -1 gpurun_out/ncu_zgemm.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sphere_project_real -s 2 -c 4 -o gpurun_out/r02_project_cfg3 $B --nband 512 > gpurun_out/ncu_proj.log 2>&1
tail -1 gpurun_out/ncu_proj.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_pass -s 3 -c 3 -o gpurun_out/r02_fft_cfg3 $B --nband 512 > gpurun_out/ncu_fft.log 2>&1
tail -1 gpurun_out/ncu_fft.log | cut -c1-200
timeout 600 ncu --set full --clock-control none -k regex:"fft_pass|sphere_project_real|zgemm_abh_kernel<float2" -s 6 -c 5 -o gpurun_out/r02_cfg2_kernels python bench.py --config cfg2 --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_cfg2.log 2>&1
tail -1 gpurun_out/ncu_cfg2.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
# e2e of the default workload after the read-order fix, with the host profile
PAWB200_PROFILE=1 timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu --no-secondary > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_g.json").read().strip().splitlines()[-1])
print("cfg3 N=1", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["checksum"])
PY
