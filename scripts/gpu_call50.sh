for t in 4 8 12 16; do echo "--- upload threads $t"; GPINV_UPLOAD_THREADS=$t timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'])"; done
