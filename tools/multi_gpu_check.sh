set -x
N=${1:-2}
python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 300 --warmup 5 2>&1 | tail -1 | cut -c1-1500
python bench.py --gpus 1 --steps 300 --warmup 5 2>&1 | tail -1 | cut -c1-400
