mkdir -p gpurun_out
echo "--- interleaved (default)"; timeout 300 python tests/gpu_cholinv_time.py 4096 2>&1 | tail -6
echo "--- behind the factorisation"; GPINV_NO_INTERLEAVED_INVERSE=1 timeout 300 python tests/gpu_cholinv_time.py 4096 2>&1 | tail -6
echo "--- n=2048, 1100"; timeout 300 python tests/gpu_cholinv_time.py 2048 2>&1 | tail -5; timeout 300 python tests/gpu_cholinv_time.py 1100 2>&1 | tail -5 | head -1
for cap in 32 100; do echo "--- cap $cap"; GPINV_INV_CAP=$cap timeout 300 python tests/gpu_cholinv_time.py 4096 2>&1 | tail -2; done
for from in 3 9; do echo "--- from $from"; GPINV_INV_FROM=$from timeout 300 python tests/gpu_cholinv_time.py 4096 2>&1 | tail -2; done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r2w_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['config']['neg_elbo'])"
