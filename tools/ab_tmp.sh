P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["stages"]["query_ms_per_step"]); print(json.dumps(d.get("db_size_sweep",{}).get("top10"))); print(json.dumps(d.get("db_size_sweep",{}).get("top10_large_batches")))'
timeout 500 python -m pytest tests/test_gpu_parity.py -x -q -k "tiled or retrieval or replay" 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>>gpurun_out/exh_err.log | python -c "$P"
