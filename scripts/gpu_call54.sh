mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
GPINV_SHARD_TRACE=20 timeout 400 $TR --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r3c_bench_n8.json 2> gpurun_out/r3c_bench_n8.err; echo "bench n8 exit $?"; grep shard-trace gpurun_out/r3c_bench_n8.err | sort | head -8
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
CUDA_VISIBLE_DEVICES=0,1,2,3 GPINV_SHARD_TRACE=20 timeout 400 $TR4 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r3c_bench_n4.json 2> gpurun_out/r3c_bench_n4.err; echo "bench n4 exit $?"; grep shard-trace gpurun_out/r3c_bench_n4.err | sort | head -4
python -c "
import json
for n in (4,8):
    d=json.load(open('gpurun_out/r3c_bench_n%d.json'%n)); print(n, d['value'], d['e2e']['value'], d['config']['neg_elbo'])"
