#!/bin/bash
# One gpurun job.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_d.log 2>&1
tail -3 gpurun_out/pytest_d.log
for f in 0 1; do
  PAWB200_FFT_FUSED=$f timeout 300 python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg2_fused$f.json 2> gpurun_out/bench_cfg2_fused$f.err
  tail -c 300 gpurun_out/bench_cfg2_fused$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_cfg2_fused$f.json"))
print("fused=$f", d["ms_per_step"], d["stage_ms_per_step"])
PY
done
timeout 900 python bench.py --steps 3 --warmup 1 --no-secondary > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
tail -c 1500 gpurun_out/bench_d.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_d.json"))
print("cfg3", d["ms_per_step"], d["stage_ms_per_step"], d["parity"])
PY
