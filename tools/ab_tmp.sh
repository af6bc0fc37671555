set -x
timeout 300 python bench.py > gpurun_out/bench_r1_final.json 2>gpurun_out/bench_err.log; tail -c 600 gpurun_out/bench_r1_final.json
timeout 200 python tools/bench_exhaustive.py --queries 64 --cpu 2000 > gpurun_out/exhaustive_100k_r1.json 2>>gpurun_out/bench_err.log
timeout 300 python tools/bench_configs.py 1 3 5 > gpurun_out/configs_r1.json 2>>gpurun_out/bench_err.log
timeout 100 python bench.py --steps 2 --warmup 3 --no-e2e --no-sweep --no-cpu-baseline > /dev/null 2>>gpurun_out/bench_err.log && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-sweep --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_build_tma -s 3 -c 1 -f -o gpurun_out/prof_build_final python bench.py --steps 1 --warmup 3 --no-e2e --no-sweep --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
timeout 100 python tools/bench_exhaustive.py --queries 8 > /dev/null 2>>gpurun_out/bench_err.log && timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_exh_screen -s 4 -c 1 -f -o gpurun_out/prof_exh_final python tools/bench_exhaustive.py --queries 8 > gpurun_out/ncu3.log 2>&1
