mkdir -p gpurun_out
timeout 600 python tests/gpu_selfcheck.py > gpurun_out/r2h_selfcheck.log 2>&1; echo "selfcheck exit $?"; grep -E "FAIL|elapsed" gpurun_out/r2h_selfcheck.log | head -5
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2h_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2h_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2h_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2h_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['config']['neg_elbo'])"
python tests/gpu_eval_timeline.py cfg5 > gpurun_out/r2h_timeline.log 2>&1; tail -2 gpurun_out/r2h_timeline.log
python - <<'PY'
import torch, time
from gpinv_b200 import ops
dev=torch.device('cuda')
n,S=4096,1024
g=torch.Generator(device=dev).manual_seed(0)
Lc=torch.tril(torch.randn(n,n,dtype=torch.float64,device=dev,generator=g))/64
Q=torch.tril(torch.randn(1,n,n,dtype=torch.float64,device=dev,generator=g))/64
E=torch.randn(1,S,n,dtype=torch.float64,device=dev,generator=g)
sv=torch.ones(1,dtype=torch.float64,device=dev); qm=torch.zeros(1,n,dtype=torch.float64,device=dev)
for we in (False, True):
    for _ in range(3): ops.stvgp_sample(Lc,sv,Q,qm,qm,E,False,we)
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): ops.stvgp_sample(Lc,sv,Q,qm,qm,E,False,we)
    b.record(); torch.cuda.synchronize(); print("stvgp_sample want_exp=%s: %.3f ms"%(we,a.elapsed_time(b)/10))
PY
