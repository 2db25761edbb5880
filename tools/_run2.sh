python -m pytest tests/test_gpu_production.py -x -q 2>&1 | tail -15
python tools/fused_vs_persistent.py > gpurun_out/r02_fused_vs_persistent.log 2>&1; cat gpurun_out/r02_fused_vs_persistent.log
for l in "" "2 256 4"; do python tools/profile_walk.py 1e6 4 spectral $l; done
python tools/profile_walk.py 1e7 3; python tools/profile_walk.py 1e7 2 const-vis
python bench.py --steps 6 --warmup 3 --no-cpu 2>/dev/null | tail -1 > gpurun_out/r02_bench_v3.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v3.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['isolated_call_ms'], d['roofline']['isolated_frac'])"
