#!/bin/bash
# N-rank bench lines of round 2 (config 3 strong scaling, config 2 band blocks), per-rank host profile in logs_n$N.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
PAWB200_PROFILE=1 timeout 900 $TR --redirects 2 --log-dir gpurun_out/logs_n$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_n$N.json 2> gpurun_out/bench_cfg3_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_cfg3_n$N.json").read().strip().splitlines()[-1])
    print("cfg3 N=$N", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["value"], d["e2e"]["value"], d["checksum"])
    for r in d.get("per_rank_ms_per_step", [])[:2]: print("  ", r)
except Exception as e: print("cfg3 N=$N failed", e)
PY
timeout 600 $TR bench.py --gpus $N --config cfg2 --steps 10 --warmup 3 > gpurun_out/r02_bench_cfg2_bands_n$N.json 2> gpurun_out/bench_cfg2_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_cfg2_bands_n$N.json").read().strip().splitlines()[-1])
    print("cfg2 N=$N (band blocks)", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["value"], d["e2e"]["value"], d["checksum"])
except Exception as e: print("cfg2 N=$N failed", e)
PY
timeout 300 $TR scripts/h2d_ceiling.py > gpurun_out/h2d_ceiling_n$N.json 2> gpurun_out/h2d_ceiling_n$N.err
cat gpurun_out/h2d_ceiling_n$N.json | cut -c1-300
du -sh gpurun_out
