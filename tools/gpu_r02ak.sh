#!/bin/bash
O=gpurun_out/r02ak
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generic_factorisation" > $O/pytest.log 2>&1; tail -15 $O/pytest.log
for n in 1000 1500 2000; do
  timeout 400 python bench.py --workload c5 --features $n --steps 20 --warmup 5 --filter-warm 10 --no-cpu-baseline > $O/bench_c5_$n.json 2> $O/bench_c5_$n.err
  python - $O/bench_c5_$n.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(d['config']['features'], 'fps', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'k', d['config']['update_rows_last_frame'], 'frac', round(d['roofline']['frac'],3), 'status', d['config']['last_frame']['status'])
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
  tail -1 $O/bench_c5_$n.err | cut -c1-300
done
timeout 300 python tools/quick_time.py 1280 720 2000 1 16 > $O/quick_c5_2000.txt 2>&1; tail -1 $O/quick_c5_2000.txt | cut -c1-400
