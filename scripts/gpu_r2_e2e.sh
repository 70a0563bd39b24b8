#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 900 python bench.py --no-kernels --no-realtime > gpurun_out/e2e_n1.json 2> gpurun_out/e2e_n1.err; tail -2 gpurun_out/e2e_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/e2e_n1.json').read().strip().splitlines()[-1]); print(1, d['ms_per_step'], d['value'], json.dumps(d['e2e'])[:900])"
if [ $NG -ge 2 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 40 --warmup 3 --no-kernels --no-realtime > gpurun_out/e2e_n2.json 2> gpurun_out/e2e_n2.err; tail -2 gpurun_out/e2e_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/e2e_n2.json').read().strip().splitlines()[-1]); print(2, d['ms_per_step'], d['value'], json.dumps(d['e2e'])[:900], d['config']['mix_abs_sum'])"
fi
