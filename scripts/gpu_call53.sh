mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29511 tests/gpu_dist_check.py cfg5 host > gpurun_out/r3b_dist_host_n8.log 2>&1; echo "dist host check exit $?"; grep -E "ok |FAIL|ms per|config|Error|error" gpurun_out/r3b_dist_host_n8.log | tail -12
timeout 400 $TR --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r3b_bench_n8.json 2> gpurun_out/r3b_bench_n8.err; echo "bench n8 exit $?"; tail -2 gpurun_out/r3b_bench_n8.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r3b_bench_n8.json')); print(d['value'], d['e2e'], d['config']['neg_elbo'])"
