#!/bin/bash
# One gpurun job.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_e.log 2>&1
tail -3 gpurun_out/pytest_e.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg2_$name.json 2> gpurun_out/bench_cfg2_$name.err
  tail -c 300 gpurun_out/bench_cfg2_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_cfg2_$name.json"))
    s=d["stage_ms_per_step"]
    print("$name", round(d["ms_per_step"],2), "fft", round(s["fft_ms"],2), "proj", round(s["project_ms"],2), "gemm", round(s["gemm_pseudo_ms"],2), "e2e", round(d["e2e"]["ms_per_step"],2))
except Exception as e: print("$name failed", e)
PY
}
run base PAWB200_FFT_TMA=0
run fused3 PAWB200_FFT_FUSED=1
run fused3_ring64 PAWB200_FFT_FUSED=1 PAWB200_FFT_RING_BYTES=67108864
run tma_sgl2 PAWB200_FFT_TMA=1 PAWB200_TMA_DBL=0 PAWB200_TMA_STAGES=2
run tma_sgl2_zl2 PAWB200_FFT_TMA=1 PAWB200_TMA_DBL=0 PAWB200_TMA_STAGES=2 PAWB200_TMA_ZL=2
run tma_sgl3_zl2 PAWB200_FFT_TMA=1 PAWB200_TMA_DBL=0 PAWB200_TMA_STAGES=3 PAWB200_TMA_ZL=2
# cuFFT-free check: launch list of the aug_recip + real-space tests
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_recip_realspace.csv \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -k "recip or realspace" > gpurun_out/ncu_recip.log 2>&1
tail -2 gpurun_out/ncu_recip.log
grep -c -i "cufft\|regular_fft\|vector_fft" gpurun_out/launches_recip_realspace.csv
timeout 600 python bench.py --config cfg4 --steps 5 --warmup 2 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
tail -c 600 gpurun_out/bench_cfg4.err; head -c 600 gpurun_out/bench_cfg4.json; echo
timeout 900 python bench.py --config cfg5 --steps 3 --warmup 1 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
tail -c 600 gpurun_out/bench_cfg5.err; head -c 600 gpurun_out/bench_cfg5.json; echo
