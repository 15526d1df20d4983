mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fused or timer_accumulates" 2>&1 | tail -15 ) > gpurun_out/fused_tests_r2g.log 2>&1
timeout 300 python scripts/fused_sweep.py > gpurun_out/fused_sweep_r2g.txt 2>&1
timeout 120 python scripts/fused_sweep.py 256 256 256 double r2c quick >> gpurun_out/fused_sweep_r2g.txt 2>&1
timeout 120 python scripts/fused_sweep.py 512 512 512 single r2c quick >> gpurun_out/fused_sweep_r2g.txt 2>&1
timeout 120 python scripts/fused_sweep.py 512 512 512 double c2c quick >> gpurun_out/fused_sweep_r2g.txt 2>&1
timeout 200 python scripts/fused_sweep.py 1024 1024 1024 single r2c quick >> gpurun_out/fused_sweep_r2g.txt 2>&1
cat gpurun_out/fused_tests_r2g.log; cat gpurun_out/fused_sweep_r2g.txt
