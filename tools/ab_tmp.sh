P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["stages"]["build_ms_per_step"], d["stages"]["query_ms_per_step"], d["roofline"]["frac"]); print(json.dumps(d.get("db_size_sweep",{}).get("top10")))'
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>>gpurun_out/exh_err.log | python -c "$P"
timeout 120 python tools/bench_configs.py 3 2>>gpurun_out/exh_err.log | cut -c1-330
