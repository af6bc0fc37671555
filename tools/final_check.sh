#!/bin/bash
# round-end style check on one B200: smoke, GPU parity tests, both bench arms
set -x
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --impl reference > gpurun_out/bench_r1_reference.json 2>gpurun_out/bench_ref_err.log; tail -c 700 gpurun_out/bench_r1_reference.json
timeout 400 python bench.py > gpurun_out/bench_r1_final.json 2>gpurun_out/bench_err.log; tail -c 300 gpurun_out/bench_r1_final.json; tail -2 gpurun_out/bench_err.log
