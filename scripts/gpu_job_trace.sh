#!/bin/bash
# Device trace of the end-to-end step at N ranks (PAWB200_TRACE marks + host section profile), per-rank logs.
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
PAWB200_TRACE=1 PAWB200_PROFILE=1 timeout 600 $TR --redirects 2 --log-dir gpurun_out/trace_n$N bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_trace_n$N.json 2> gpurun_out/bench_trace_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_trace_n$N.json").read().strip().splitlines()[-1])
print("traced cfg3 N=$N", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
PY
du -sh gpurun_out
