#!/bin/bash
bash scripts/gpu_job_multi.sh 8
bash scripts/gpu_job_multi.sh 4
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1; nproc >> gpurun_out/topo8.txt; free -g >> gpurun_out/topo8.txt
