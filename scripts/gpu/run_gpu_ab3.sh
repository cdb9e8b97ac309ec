#!/bin/bash
mkdir -p gpurun_out
CFG=c3 ./run_gpu_ab2.sh "$@"
timeout 300 python scripts/wait_profile.py c3p > gpurun_out/wait_profile.log 2>&1; cat gpurun_out/wait_profile.log
