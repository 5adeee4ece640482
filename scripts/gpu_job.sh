#!/bin/bash
# One gpurun job.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_f.log 2>&1
tail -3 gpurun_out/pytest_f.log
PAWB200_KEEP_BOXES_BYTES=0 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_f_lazy.log 2>&1
tail -3 gpurun_out/pytest_f_lazy.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_recip_realspace.csv \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -k "recip or realspace or fft_check" > gpurun_out/ncu_recip.log 2>&1
tail -2 gpurun_out/ncu_recip.log
echo "cuFFT launches:" $(grep -c -i "cufft\|regular_fft\|vector_fft" gpurun_out/launches_recip_realspace.csv)
timeout 1200 python bench.py --steps 5 --warmup 2 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err
tail -c 1500 gpurun_out/bench_f.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_f.json"))
print("cfg3", d["ms_per_step"], d["e2e"]["ms_per_step"], d["stage_ms_per_step"], d["parity"]["ok"], d["cpu_baseline"]["value"])
c=d["cfg2"]; print("cfg2", c["ms_per_step"], c["e2e"]["ms_per_step"], c["stage_ms_per_step"], c["parity"]["ok"], c["cpu_baseline"]["value"])
PY
