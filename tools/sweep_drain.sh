#!/bin/bash
# bench value / e2e / isolated launch time vs the drain-phase hand-over threshold (MC3D_DRAIN_GIVE; unset = automatic)
cd "$(dirname "$0")/.."
for g in auto 0 8 12 16 20 24 31; do
    if [ "$g" = auto ]; then unset MC3D_DRAIN_GIVE; else export MC3D_DRAIN_GIVE=$g; fi
    timeout 300 python bench.py --steps 400 --warmup 20 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('give %-5s value %.4e e2e %.4e iso_ms %.3f frac %.3f' % ('$g', d['value'], d['e2e']['value'], d['roofline']['isolated_launch_ms'], d['roofline']['frac']))"
done
