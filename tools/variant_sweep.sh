#!/bin/bash
# single-GPU sweep over the kernel table (vrenb200_sort_config::variant) and the tile-id modes; one bench line each
for v in 2 3 5 6 8 9 11 12; do
  python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline --variant $v 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('variant $v', d['details']['variant'], 'Gpairs/s %.2f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'pass_ms %.3f' % d['roofline']['kernel_ms'], 'frac %.3f' % d['roofline']['frac'])
    elif 'rror' in l: print(l.strip())
"
done
for t in block ticket; do for r in verified sampled unverified match; do
  python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline --ranking $r --tile-ids $t 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ranking $r tile-ids $t', 'Gpairs/s %.2f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'pass_ms %.3f' % d['roofline']['kernel_ms'], 'frac %.3f' % d['roofline']['frac'], 'violations', d['details']['ranking_check_failures'])
    elif 'rror' in l: print(l.strip())
"
done; done
