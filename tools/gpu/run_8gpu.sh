#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "rc=$?"; tail -c 600 gpurun_out/r02_bench_n8.err
python - <<'P'
import json
for line in open('gpurun_out/r02_bench_n8.json'):
    line=line.strip()
    if line.startswith('{'):
        b=json.loads(line); print({k:b[k] for k in ('value','ms_per_step','n_gpus')}); print(json.dumps(b['sharded_c3'], indent=1))
P
