#!/bin/bash
# development aid: bench.py at several GPU counts on one box (the driver runs the same sweep at round end)
#   tools/scale_sweep.sh WORKLOAD STEPS "1 2 4 8"
mkdir -p gpurun_out
W=${1:-sp13_2p24}; STEPS=${2:-20}; NS=${3:-"1 2 4 8"}
for n in $NS; do
  if [ $n -eq 1 ]; then
    python bench.py --workload $W --steps $STEPS --no-cpu-baseline > gpurun_out/scale_${W}_$n.log 2>&1
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --workload $W --steps $STEPS > gpurun_out/scale_${W}_$n.log 2>&1
  fi
  tail -1 gpurun_out/scale_${W}_$n.log | python -c "
import sys, json
l = sys.stdin.readline()
try:
    d = json.loads(l); print(d['n_gpus'], 'value %.4e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel_ms %.4f' % d['roofline']['kernel_ms'], 'e2e %.3e' % d['e2e']['value'], d['config']['parallelism'][:60], d['clocks'])
except Exception as e:
    print('FAILED', l[:300])
"
done
