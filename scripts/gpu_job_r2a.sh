#!/bin/bash
# Round-2 evidence job: GPU tests, default bench, ncu launch lists and full-set captures of the dominant kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_a.log 2>&1
tail -3 gpurun_out/pytest_a.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
tail -c 800 gpurun_out/bench_a.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_a.json").read().strip().splitlines()[-1])
print("cfg3", d["ms_per_step"], d["e2e"]["ms_per_step"], d["stage_ms_per_step"], d["parity"]["ok"], d["cpu_baseline"]["value"])
c=d["cfg2"]; print("cfg2", c["ms_per_step"], c["e2e"]["ms_per_step"], c["stage_ms_per_step"], c["parity"]["ok"], c["cpu_baseline"]["value"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_recip_realspace.csv \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -k "recip or realspace or fft_check" > gpurun_out/ncu_recip.log 2>&1
tail -2 gpurun_out/ncu_recip.log
echo "cuFFT launches:" $(grep -c -i "cufft\|regular_fft\|vector_fft" gpurun_out/r02_launches_recip_realspace.csv)
B="python bench.py --steps 1 --warmup 0 --no-cpu --no-secondary"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_cfg3.csv $B > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zgemm_abh_kernel -s 1 -c 2 -o gpurun_out/r02_zgemm_cfg3 $B > gpurun_out/ncu_zgemm.log 2>&1
tail -1 gpurun_out/ncu_zgemm.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sphere_project_real -s 2 -c 4 -o gpurun_out/r02_project_cfg3 $B --nband 512 > gpurun_out/ncu_proj.log 2>&1
tail -1 gpurun_out/ncu_proj.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_pass -s 3 -c 3 -o gpurun_out/r02_fft_cfg3 $B --nband 512 > gpurun_out/ncu_fft.log 2>&1
tail -1 gpurun_out/ncu_fft.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fft_pass|sphere_project_real|zgemm_abh_kernel<float2" -s 6 -c 5 -o gpurun_out/r02_cfg2_kernels python bench.py --config cfg2 --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_cfg2.log 2>&1
tail -1 gpurun_out/ncu_cfg2.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
