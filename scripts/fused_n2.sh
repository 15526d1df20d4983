#!/bin/bash
# fused forward Z + peer-storing Y on a slab grid: bench.py under torchrun with ACCFFT_B200_FUSED = 0 / 2 / 1.  usage: scripts/fused_n2.sh N TAG
N=${1:-2}; TAG=${2:-r2q}
mkdir -p gpurun_out
for F in ${MODES:-0 2 1}; do
ACCFFT_B200_FUSED=$F timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+F)) \
  bench.py --gpus $N --steps 30 --warmup 5 --no-configs --no-e2e --no-parity 2>/dev/null | python -c "
import sys,json
L=[l for l in sys.stdin.read().splitlines() if l.startswith('{')]
d=json.loads(L[-1])
print('N=$N FUSED=$F ms %.4f value %.0f frac %.3f rt %.1e' % (d['ms_per_step'], d['value'], d['step_roofline']['frac_of_roofline'], d['config']['round_trip_rel_l2']), {k:round(v['ms']*v['launches_per_step'],4) for k,v in d['stages'].items()}, d.get('fused_pair'))
" | tee -a gpurun_out/fused_n${N}_$TAG.txt
done
