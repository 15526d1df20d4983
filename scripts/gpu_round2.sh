#!/bin/bash
# One GPU-box visit (1 GPU), round 2: tests, smoke, bench (both arms), stage profiles, SASS evidence, ncu launch list + full
# captures of the dominant kernels.  Outputs in gpurun_out/.   usage: scripts/gpu_round2.sh TAG [SKIP_TESTS=1]
mkdir -p gpurun_out
TAG=${1:-r2}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi_$TAG.txt
nproc >> gpurun_out/smi_$TAG.txt
if [ -z "$SKIP_TESTS" ]; then
( time timeout 1500 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -25 ) > gpurun_out/gpu_tests_$TAG.log 2>&1
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
timeout 600 python scripts/stage_profile.py > gpurun_out/stage_profile_$TAG.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_$TAG.log 2>&1
if [ -z "$SKIP_FULL" ]; then
# double: one X-stage launch (plain strided kernel), one Y-stage launch (TMA kernel), the two Z kernels
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_strided_c2c -s 6 -c 1 -f -o gpurun_out/prof_strided_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full1_$TAG.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_strided_tma -s 6 -c 1 -f -o gpurun_out/prof_tma_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full2_$TAG.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_contig -s 6 -c 2 -f -o gpurun_out/prof_contig_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full3_$TAG.log 2>&1
# single: the lane-pair strided kernel at 512^3
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_strided_c2c_pair -s 8 -c 1 -f -o gpurun_out/prof_pair_$TAG \
    python scripts/stage_profile.py 512 512 512 single r2c > gpurun_out/ncu_full4_$TAG.log 2>&1
python scripts/ncu_summary.py gpurun_out/ncu_summary_$TAG.txt gpurun_out/prof_strided_$TAG.ncu-rep gpurun_out/prof_tma_$TAG.ncu-rep \
    gpurun_out/prof_contig_$TAG.ncu-rep gpurun_out/prof_pair_$TAG.ncu-rep > /dev/null 2>&1
fi
# keep the merged directory under the 64 MiB limit: drop the biggest reports first
while [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; do big=$(ls -S gpurun_out/*.ncu-rep 2>/dev/null | head -1); [ -z "$big" ] && break; rm -f "$big"; done
du -sh gpurun_out; ls -la gpurun_out | tail -30
