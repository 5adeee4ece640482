#!/bin/bash
# Block-ordered deferred ingest: tests (1 GPU part), then N-rank e2e numbers and a traced short run.
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_h.log 2>&1
tail -3 gpurun_out/pytest_h.log
timeout 600 python bench.py --nband 512 --steps 3 --warmup 3 --no-cpu --no-secondary 2> gpurun_out/b3.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3-512 N=1', 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['checksum'])"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_n$N.json 2> gpurun_out/bench_cfg3_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_cfg3_n$N.json").read().strip().splitlines()[-1])
    print("cfg3 N=$N", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["value"], d["e2e"]["value"], d["checksum"])
except Exception as e: print("cfg3 N=$N failed", e)
PY
PAWB200_TRACE=1 timeout 600 $TR --redirects 2 --log-dir gpurun_out/trace_n$N bench.py --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_trace_n$N.json 2> gpurun_out/bench_trace_n$N.err
du -sh gpurun_out
