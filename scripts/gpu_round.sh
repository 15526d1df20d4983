#!/bin/bash
# One GPU-box visit (1 GPU): tests, smoke, bench (both arms), ncu launch list + full captures.  Outputs in gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-r01}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi_$TAG.txt
nproc >> gpurun_out/smi_$TAG.txt
( time timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
timeout 600 python scripts/cfg45_local.py all > gpurun_out/cfg45_$TAG.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_$TAG.log 2>&1
if [ -z "$SKIP_FULL" ]; then
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_strided -s 12 -c 1 -f -o gpurun_out/prof_strided_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full1_$TAG.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_contig -s 6 -c 2 -f -o gpurun_out/prof_contig_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full2_$TAG.log 2>&1
python scripts/ncu_summary.py gpurun_out/ncu_summary_$TAG.txt gpurun_out/prof_strided_$TAG.ncu-rep gpurun_out/prof_contig_$TAG.ncu-rep > /dev/null 2>&1
fi
# keep the merged directory under the 64 MiB limit: drop the biggest reports first
while [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; do big=$(ls -S gpurun_out/*.ncu-rep 2>/dev/null | head -1); [ -z "$big" ] && break; rm -f "$big"; done
du -sh gpurun_out; ls -la gpurun_out
