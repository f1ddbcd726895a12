mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 tests/gpu_dist_check.py cfg5 > gpurun_out/r3a_dist_cfg5_n2.log 2>&1; echo "dist check exit $?"; grep -E "ok |FAIL|ms per|config|Error|error" gpurun_out/r3a_dist_cfg5_n2.log | tail -30
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r3a_bench_n2.json 2> gpurun_out/r3a_bench_n2.err; echo "bench n2 exit $?"; tail -2 gpurun_out/r3a_bench_n2.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r3a_bench_n2.json')); print(d['value'], d['e2e'], d['config']['neg_elbo'])"
