for inflight in 1 2 3 4; do for launch in 4,256,4 8,128,4 8,128,2; do
echo "inflight $inflight launch $launch"; python bench.py --steps 300 --warmup 5 --inflight $inflight --launch $launch | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  value %.4e ms/step %.3f e2e %.4e e2e_ms %.3f iso_ms %.3f frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['isolated_launch_ms'], d['roofline']['frac']))"
done; done
