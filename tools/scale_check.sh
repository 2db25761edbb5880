# 1 -> N GPU scaling of bench.py (torchrun, one rank per GPU) + config C5 on all GPUs + multi-GPU tests
NMAX=${1:-8}
python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2
for N in 1 2 4 8; do
  if [ $N -le $NMAX ]; then
    if [ $N -eq 1 ]; then python bench.py --gpus 1 --steps 500 --warmup 10 > gpurun_out/scale_n$N.json 2>gpurun_out/scale_n$N.err
    else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 500 --warmup 10 > gpurun_out/scale_n$N.json 2>gpurun_out/scale_n$N.err; fi
    tail -1 gpurun_out/scale_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=%d value %.4e photons/s  events/s %.4e  ms/step %.3f  e2e %.4e  frac %.3f' % (d['n_gpus'], d['value'], d['events_per_s'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))"
  fi
done
python tools/run_configs.py c5
