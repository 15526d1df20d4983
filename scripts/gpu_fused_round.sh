#!/bin/bash
# One GPU-box visit for the fused local pair: its tests, the parameter sweep, stage profile, one ncu --set full capture.  usage: scripts/gpu_fused_round.sh TAG
TAG=${1:-r2i}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fused or timer_accumulates or (r2c_forward and tuned) or (c2c_against and tuned) or operators_against or single_axis_transforms" 2>&1 | tail -15 ) > gpurun_out/fused_tests_$TAG.log 2>&1
timeout 300 python scripts/fused_sweep.py 256 256 256 double r2c quick > gpurun_out/fused_sweep_$TAG.txt 2>&1
timeout 300 python scripts/fused_sweep.py >> gpurun_out/fused_sweep_$TAG.txt 2>&1
timeout 120 python scripts/fused_sweep.py 512 512 512 single r2c quick >> gpurun_out/fused_sweep_$TAG.txt 2>&1
timeout 120 python scripts/fused_sweep.py 512 512 512 double c2c quick >> gpurun_out/fused_sweep_$TAG.txt 2>&1
ACCFFT_B200_FUSED=0 timeout 600 python scripts/stage_profile.py > gpurun_out/stage_profile_$TAG.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 4 -c 2 -f -o gpurun_out/prof_fused_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fused_$TAG.log 2>&1
python scripts/ncu_summary.py gpurun_out/ncu_summary_fused_$TAG.txt gpurun_out/prof_fused_$TAG.ncu-rep > /dev/null 2>&1
cat gpurun_out/fused_tests_$TAG.log; cat gpurun_out/fused_sweep_$TAG.txt; cat gpurun_out/stage_profile_$TAG.txt; cut -c1-200 gpurun_out/ncu_summary_fused_$TAG.txt
