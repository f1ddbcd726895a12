timeout 600 python tests/gpu_pipeline_spread.py 640 2>&1 | tail -4
