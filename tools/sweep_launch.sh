#!/bin/bash
# bench value vs launch shape (blocks_per_sm,block_threads,refill_threshold) with 8 steps in flight
cd "$(dirname "$0")/.."
for l in "" "1,128,4" "2,128,4" "3,128,4" "1,512,4" "1,256,3" "1,256,6" "2,256,4"; do
    timeout 300 python bench.py --steps 500 --warmup 20 ${l:+--launch $l} 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('launch %-10s value %.4e e2e %.4e iso_ms %.3f frac %.3f grid %d' % ('${l:-auto}', d['value'], d['e2e']['value'], d['roofline']['isolated_launch_ms'], d['roofline']['frac'], d['config']['grid_blocks']))"
done
