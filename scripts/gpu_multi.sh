#!/bin/bash
# Multi-GPU parity and bench on N GPUs of one box:  gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_multi.sh N'
# (never under ncu).  Parity: every rank's result against the oracle fixture of cfg5 and the longdouble arbiter.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 tests/gpu_dist_check.py cfg5 > gpurun_out/dist_check_n$N.log 2>&1; echo "dist check exit $?"
grep -E "ok |FAIL|ms per|config|Error" gpurun_out/dist_check_n$N.log | tail -30
timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print(d['value'], d['e2e'], d['config']['neg_elbo'])"
