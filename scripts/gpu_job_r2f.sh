#!/bin/bash
# Round-2 evidence on one B200: default bench line, reference arm, aux configs, ncu launch lists + full-set captures
# (exported to csv on the box: gpurun copies back 64 MiB at most), compute-sanitizer on the kernels changed this round.
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_n1.json 2> gpurun_out/bench_n1.err
tail -c 400 gpurun_out/bench_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_cfg3_n1.json").read().strip().splitlines()[-1])
print("cfg3", d["ms_per_step"], d["e2e"]["ms_per_step"], {k:round(v,1) for k,v in d["stage_ms_per_step"].items()}, d["parity"]["ok"], d["cpu_baseline"]["value"], {k:round(v["frac"],3) for k,v in d["kernels"].items()}, d["roofline"]["frac_issued"])
c=d["cfg2"]; print("cfg2", c["ms_per_step"], c["e2e"]["ms_per_step"], {k:round(v,2) for k,v in c["stage_ms_per_step"].items()}, c["parity"]["ok"], c["cpu_baseline"]["value"], {k:round(v["frac"],3) for k,v in c["kernels"].items()})
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/bench_ref.err
cut -c1-600 gpurun_out/r02_bench_reference_arm.json
timeout 600 python bench.py --config cfg4 > gpurun_out/r02_bench_cfg4.json 2> gpurun_out/bench_cfg4.err; cut -c1-400 gpurun_out/r02_bench_cfg4.json
timeout 600 python bench.py --config cfg5 > gpurun_out/r02_bench_cfg5.json 2> gpurun_out/bench_cfg5.err; cut -c1-400 gpurun_out/r02_bench_cfg5.json
B="python bench.py --steps 1 --warmup 0 --no-cpu --no-secondary"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_cfg3.csv $B > gpurun_out/ncu_launches.log 2>&1
gzip -f gpurun_out/r02_launches_cfg3.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_recip_realspace.csv \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -k "recip or realspace or fft_check" > gpurun_out/ncu_recip.log 2>&1
tail -2 gpurun_out/ncu_recip.log
echo "cuFFT launches:" $(grep -c -i "cufft\|regular_fft\|vector_fft" gpurun_out/r02_launches_recip_realspace.csv)
gzip -f gpurun_out/r02_launches_recip_realspace.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"zgemm_abh_kernel<float2" -c 1 -o gpurun_out/r02_ncu_zgemm_cfg3 $B > gpurun_out/ncu_zgemm.log 2>&1
bash scripts/ncu_export.sh gpurun_out/r02_ncu_zgemm_cfg3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sphere_project_real -s 2 -c 3 -o gpurun_out/r02_ncu_project_cfg3 $B --nband 512 > gpurun_out/ncu_proj.log 2>&1
bash scripts/ncu_export.sh gpurun_out/r02_ncu_project_cfg3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_pass -s 3 -c 3 -o gpurun_out/r02_ncu_fft_cfg3 $B --nband 512 > gpurun_out/ncu_fft.log 2>&1
bash scripts/ncu_export.sh gpurun_out/r02_ncu_fft_cfg3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fft_pass|sphere_project_real|zgemm_abh_kernel<float2" -s 6 -c 5 -o gpurun_out/r02_ncu_cfg2_kernels python bench.py --config cfg2 --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_cfg2.log 2>&1
bash scripts/ncu_export.sh gpurun_out/r02_ncu_cfg2_kernels
K="ga4_projections or synthetic_golden or awkward or 17_bands or aug_recip_vs or realspace_projection"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fft_paths.py -m gpu -x -q -k "$K" > gpurun_out/r02_memcheck.log 2>&1
tail -4 gpurun_out/r02_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fft_paths.py -m gpu -x -q -k "ga4_projections or awkward or 17_bands or synthetic_golden" > gpurun_out/r02_racecheck.log 2>&1
tail -4 gpurun_out/r02_racecheck.log
du -sh gpurun_out
