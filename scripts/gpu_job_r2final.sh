#!/bin/bash
# Final 1-GPU record of round 2: full GPU suite (default and deferred-projection mode), default bench, reference arm.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1
tail -2 gpurun_out/pytest_final.log
PAWB200_KEEP_BOXES_BYTES=0 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final_lazy.log 2>&1
tail -2 gpurun_out/pytest_final_lazy.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_n1.json 2> gpurun_out/bench_n1.err
tail -c 300 gpurun_out/bench_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_cfg3_n1.json").read().strip().splitlines()[-1])
print("cfg3", d["ms_per_step"], d["e2e"]["ms_per_step"], {k:round(v,1) for k,v in d["stage_ms_per_step"].items()}, d["parity"]["ok"], d["cpu_baseline"]["value"], {k:round(v["frac"],3) for k,v in d["kernels"].items()}, d["roofline"]["frac_issued"])
c=d["cfg2"]; print("cfg2", c["ms_per_step"], c["e2e"]["ms_per_step"], {k:round(v,2) for k,v in c["stage_ms_per_step"].items()}, c["parity"]["ok"], c["cpu_baseline"]["value"], {k:round(v["frac"],3) for k,v in c["kernels"].items()})
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/bench_ref.err
cut -c1-200 gpurun_out/r02_bench_reference_arm.json
B="python bench.py --steps 1 --warmup 0 --no-cpu --no-secondary"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zgemm_abh_kernel -c 1 -o gpurun_out/r02_ncu_zgemm_cfg3 $B > gpurun_out/ncu_zgemm.log 2>&1
bash scripts/ncu_export.sh gpurun_out/r02_ncu_zgemm_cfg3
du -sh gpurun_out
