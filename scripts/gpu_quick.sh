#!/bin/bash
# quick 1-GPU check after a kernel change: a parity subset, the stage profile of the headline sizes.  usage: scripts/gpu_quick.sh TAG
TAG=${1:-q}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ref_golden.py -q -m gpu -x -k "not full_size and not wide-"  2>&1 | tail -6 ) > gpurun_out/quick_tests_$TAG.log 2>&1
timeout 300 python scripts/stage_profile.py 512 512 512 double r2c 512 512 512 single r2c 256 256 256 double r2c 1024 1024 1024 single r2c > gpurun_out/stage_profile_$TAG.txt 2>&1
cat gpurun_out/quick_tests_$TAG.log gpurun_out/stage_profile_$TAG.txt
