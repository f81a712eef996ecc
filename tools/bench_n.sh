#!/bin/bash
# usage: tools/bench_n.sh N "extra bench args" ... : one bench line per extra-args string, compact summary
N=$1; shift
for extra in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 6 --warmup 3 --no-secondary $extra 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('N=$N', '$extra', 'Gpairs/s %.2f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e %.2f' % d['e2e']['value'], 'pass_ms %.3f' % d['roofline']['kernel_ms'], d.get('phases_rank0_last_step'))
    elif 'rror' in l or 'assert' in l: print(l.strip())
"
done
