# 1 -> 8 GPU scaling of bench.py (torchrun, one rank per GPU) + multi-GPU tests + config C5 on 1/2/4/8 GPUs
python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_scale_n$N.json 2>gpurun_out/r02_scale_n$N.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_scale_n$N.json 2>gpurun_out/r02_scale_n$N.err; fi
  tail -1 gpurun_out/r02_scale_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('N=%d value %.4e photons/s  events/s %.4e  ms/step %.3f  frac %.3f | e2e %.4e  d2h %.1f GB/s per GPU of %.1f measured (link_frac %.3f) | digest %s' % (d['n_gpus'], d['value'], d['events_per_s'], d['ms_per_step'], d['roofline']['frac'], e['value'], e['d2h_GBps_per_gpu'], e['link_GBps_per_gpu'], e['link_frac'], d['invariance_digest'][:16]))"
done
python tools/run_configs.py c5
