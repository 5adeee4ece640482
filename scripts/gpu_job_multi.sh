#!/bin/bash
# Multi-GPU gpurun job: N = $1 ranks.  Outputs under gpurun_out/.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 2 > gpurun_out/bench_cfg3_n$N.json 2> gpurun_out/bench_cfg3_n$N.err
grep "^\[rank" gpurun_out/bench_cfg3_n$N.err | cut -c1-400; tail -c 600 gpurun_out/bench_cfg3_n$N.err | grep -v "^\[rank"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_cfg3_n$N.json").read().strip().splitlines()[-1])
    print("cfg3 N=$N", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["value"], d["e2e"]["value"])
except Exception as e: print("cfg3 N=$N failed", e)
PY
timeout 600 $TR bench.py --gpus $N --config cfg2 --steps 10 --warmup 3 > gpurun_out/bench_cfg2_n$N.json 2> gpurun_out/bench_cfg2_n$N.err
grep "^\[rank" gpurun_out/bench_cfg2_n$N.err | cut -c1-400; tail -c 600 gpurun_out/bench_cfg2_n$N.err | grep -v "^\[rank"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_cfg2_n$N.json").read().strip().splitlines()[-1])
    print("cfg2 N=$N (band blocks)", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["value"], d["e2e"]["value"], d["checksum"])
except Exception as e: print("cfg2 N=$N failed", e)
PY
timeout 300 $TR scripts/h2d_ceiling.py > gpurun_out/h2d_ceiling_n$N.json 2> gpurun_out/h2d_ceiling_n$N.err
cat gpurun_out/h2d_ceiling_n$N.json; tail -c 300 gpurun_out/h2d_ceiling_n$N.err
if [ "$N" = "2" ]; then
  timeout 300 python scripts/h2d_ceiling.py > gpurun_out/h2d_ceiling_n1.json 2>&1; cat gpurun_out/h2d_ceiling_n1.json
  PAWB200_KEEP_BOXES_BYTES=0 timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fft_paths.py::test_17_bands_group_tail_and_offsite_reuse_of_resident_boxes > gpurun_out/pytest_g_lazy.log 2>&1
  tail -3 gpurun_out/pytest_g_lazy.log
fi
