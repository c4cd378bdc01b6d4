TAG=r01u
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1
timeout 400 python bench.py > $O/bench_c3.json 2> $O/bench_c3.err
timeout 200 python bench.py --workload c2 > $O/bench_c2.json 2> $O/bench_c2.err
timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
timeout 300 python tools/quick_time.py 640 480 500 1 80 > $O/quick_c3.txt 2>&1
timeout 300 python tools/quick_time.py 640 480 200 256 30 > $O/quick_c4.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_bench_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --filter-warm 4 > $O/ncu_bench.log 2>&1
python tools/agg_launches.py $O/launches_bench_c3.csv > $O/launches_bench_c3.txt 2>&1
tail -3 $O/pytest_gpu.log; tail -1 $O/smoke.log; cut -c1-200 $O/bench_c3.json
