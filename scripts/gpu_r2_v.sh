#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-realtime --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 600 gpurun_out/bench_n2.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_n2.json').read())
print({k:l[k] for k in ('value','n_gpus','ms_per_step','e2e','parity','gpu_launches','clocks')}); print(l['config']['parallelism'])
PY
timeout 600 python bench.py --steps 20 --warmup 5 --no-realtime --no-cpu-baseline --no-kernels > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; tail -c 300 gpurun_out/bench_n1b.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_n1b.json').read())
print({k:l[k] for k in ('value','n_gpus','ms_per_step','e2e','parity','gpu_launches','clocks')}); print({k:l['roofline'][k] for k in ('achieved','peak','frac','peak_burst','frac_of_burst_peak','kernel_ms')})
PY
