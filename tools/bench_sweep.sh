for cfg in "4 4,256,4" "4 1,256,4" "4 2,128,4" "8 1,128,4" "8 1,256,4" "8 2,128,4" "6 1,256,4" "8 1,128,2"; do set -- $cfg
echo "inflight $1 launch $2"; python bench.py --steps 400 --warmup 8 --inflight $1 --launch $2 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  value %.4e ms/step %.3f e2e %.4e e2e_ms %.3f iso_ms %.3f frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['isolated_launch_ms'], d['roofline']['frac']))"
done
