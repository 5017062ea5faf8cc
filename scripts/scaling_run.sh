#!/bin/bash
# Weak-scaling curve of bench.py on one 8-GPU node for the BASELINE.json multi-GPU configs (run under `gpurun --gpus 8`).
# usage: scripts/scaling_run.sh out_file [workload ...]
out=${1:-gpurun_out/r2_scaling.txt}
shift
wls=${@:-cfg2_endovis18_384px_T10_7obj_x8clips cfg3_cholec_512px_T8_13obj_x1clip cfg4_1024px_T8_4obj_x1clip}
: > $out
for wl in $wls; do
  for n in 1 2 4 8; do
    if [ $n -eq 1 ]; then
      line=$(python bench.py --gpus 1 --steps 10 --warmup 3 --workload $wl --no-extras --no-cpu-baseline 2>/dev/null | tail -1)
    else
      line=$(python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 10 --warmup 3 --workload $wl --no-extras --no-cpu-baseline 2>/dev/null | tail -1)
    fi
    echo "$line" | python -c "
import json,sys
d=json.loads(sys.stdin.read())
e=d.get('e2e') or {}
print('%s N=%d value=%.1f %s ms/step=%.3f e2e=%.1f e2e_ms=%.3f' % (d['config']['workload'], d['n_gpus'], d['value'], d['unit'], d['ms_per_step'], e.get('value', 0), e.get('ms_per_step', 0)))" >> $out
  done
done
cat $out
