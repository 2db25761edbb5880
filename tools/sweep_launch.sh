#!/bin/bash
# bench value vs launch shape (blocks_per_sm,block_threads,refill_threshold; 255 = automatic grid)
cd "$(dirname "$0")/.."
for l in "" "255,256,2" "255,256,6" "255,256,8" "255,256,12" "255,256,16" "255,128,8" "255,512,8" "2,256,8"; do
    timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu ${l:+--launch $l} 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('launch %-12s value %.4e e2e %.4e iso_ms %.3f frac %.3f grid %d' % ('${l:-auto}', d['value'], d['e2e']['value'], d['roofline']['isolated_call_ms'], d['roofline']['frac'], d['run_info']['grid_blocks']))"
done
