#!/bin/bash
set -u
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02d_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02d_gpu_tests.log
timeout 600 python bench.py > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02d_bench_ref.json 2> gpurun_out/r02d_bench_ref.err; echo "ref rc=$?"
timeout 900 bash tools/profile_round.sh r02d; echo "profile rc=$?"
