set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -q -k "not cfg3b_small_full" -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -25 gpurun_out/r2a_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; echo "ref exit $?"
(python tests/gpu_arbiter_report.py; GPINV_POTRF_BWD=murray python tests/gpu_arbiter_report.py; GPINV_BWD_BLOCK=256 python tests/gpu_arbiter_report.py) > gpurun_out/r2a_arbiter.jsonl 2> gpurun_out/r2a_arbiter.err
cat gpurun_out/r2a_bench.json | head -c 3000
