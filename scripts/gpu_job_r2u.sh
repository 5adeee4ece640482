#!/bin/bash
# three-factor line transforms: the awkward-grid tests (projections, one real-space state, density per grid)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fft_paths.py -m gpu -x -q -k "awkward" > gpurun_out/pytest_u.log 2>&1
tail -15 gpurun_out/pytest_u.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fft_check or ga4_projections or aug_recip_vs" > gpurun_out/pytest_u2.log 2>&1
tail -3 gpurun_out/pytest_u2.log
