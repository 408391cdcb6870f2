#!/bin/bash
# 2-GPU box: the NCCL test of the batch-sharded logpdf and the bench with its sharded-C3 leg at N = 2
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -4
echo "== NCCL test"
timeout 600 python -m pytest tests/test_dist.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/r02_pytest_nccl.log
echo "== bench --gpus 2 (torchrun)"
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "rc=$?"; tail -c 800 gpurun_out/r02_bench_n2.err
python - <<'P'
import json
for line in open('gpurun_out/r02_bench_n2.json'):
    line=line.strip()
    if line.startswith('{'):
        b=json.loads(line); print({k:b[k] for k in ('value','ms_per_step','n_gpus')}); print(json.dumps(b['sharded_c3'], indent=1))
P
