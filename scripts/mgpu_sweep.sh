#!/bin/bash
# Exchange-mode sweep on an 8-GPU box (run through gpurun --gpus 8): direct peer stores vs copy-engine vs TMA push transfers.
# usage: scripts/mgpu_sweep.sh  -> prints one line per configuration
run() {  # N chunks lanes push
  ACCFFT_B200_CE=$2 ACCFFT_B200_CE_LANES=$3 ACCFFT_B200_PUSH=$4 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 \
    --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $1 --steps 30 --warmup 5 --no-configs --no-e2e --no-parity 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
x=d['exchange']
print('N=$1 ce=$2 lanes=$3 push=$4 ms %.4f frac %.3f rt %.1e' % (d['ms_per_step'], d['step_roofline']['frac_of_roofline'], d['config']['round_trip_rel_l2']), {k:round(v['ms']*v['launches_per_step'],3) for k,v in d['stages'].items()}, x.get('transfer_busy_ms_per_step'), x.get('nvlink_out_gbs'))
"
}
run 8 0 4 0
run 8 4 4 0
run 8 4 4 16
run 8 2 4 24
run 8 4 7 16
run 4 0 4 0
run 4 4 4 16
