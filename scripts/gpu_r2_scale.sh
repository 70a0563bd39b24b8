#!/bin/bash
# strong scaling of the bench job over the GPUs of one box: N = 2, 4, 8 as far as the box has them
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for N in 2 4 8; do
  if [ $N -le $NG ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 40 --warmup 3 --no-kernels > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
    tail -1 gpurun_out/r2_bench_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'], d['parity']['rel_l2'], d['clocks']['sm_mhz'])"
  fi
done
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -3
