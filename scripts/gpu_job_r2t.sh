#!/bin/bash
# Last checks of the round on one GPU: the ingest / prelaunch tests, the suite in deferred-projection mode, smoke().
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fft_paths.py -m gpu -x -q -k "async or chunked or sharded" > gpurun_out/pytest_t.log 2>&1
tail -3 gpurun_out/pytest_t.log
PAWB200_KEEP_BOXES_BYTES=0 timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fft_paths.py::test_17_bands_group_tail_and_offsite_reuse_of_resident_boxes > gpurun_out/pytest_t_lazy.log 2>&1
tail -3 gpurun_out/pytest_t_lazy.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
