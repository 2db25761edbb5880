#!/bin/bash
# bench value / e2e vs steps in flight (library slots)
cd "$(dirname "$0")/.."
for d in 8 10 12 16; do
    timeout 300 python bench.py --steps 600 --warmup 20 --inflight $d 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('inflight %-3s value %.4e (%.4f ms) e2e %.4e (%.4f ms)' % ('$d', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))"
done
