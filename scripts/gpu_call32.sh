mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/gpu_dist_check.py cfg5 > gpurun_out/r2l_dist_cfg5.log 2>&1; echo "dist check exit $?"; grep -vE "^W|^\*|Warning|warn" gpurun_out/r2l_dist_cfg5.log | tail -25
GPINV_SHARD_REVERSE=0 timeout 600 $TR --master-port 29512 tests/gpu_dist_check.py cfg5 > gpurun_out/r2l_dist_cfg5_repl.log 2>&1; echo "dist check (replicated) exit $?"; grep -E "ms per|FAIL|config" gpurun_out/r2l_dist_cfg5_repl.log | tail -8
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2l_bench_n2.json 2> gpurun_out/r2l_bench_n2.err; echo "bench n2 exit $?"; tail -2 gpurun_out/r2l_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r2l_bench_n2.json')); print(d['value'], d['e2e'], d['config']['neg_elbo'], d['config']['launches_per_eval'])"
