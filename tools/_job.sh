#!/bin/bash
set -u
args="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-lz --configs="
HOH_GPU_LIB=$PWD/hoh-ans_b200/csrc/variants/libhohgpu_old.so timeout 120 python bench.py $args > gpurun_out/ab_old.json 2> gpurun_out/ab_old.err
timeout 120 python bench.py $args > gpurun_out/ab_default.json 2> gpurun_out/ab_default.err
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_encode_images.py tests/test_gpu_config_sizes.py -x -q -m gpu > gpurun_out/ab_tests.log 2>&1
tail -3 gpurun_out/ab_tests.log
python - <<'PY'
import json
for t in ("old","default"):
    d=json.loads(open(f"gpurun_out/ab_{t}.json").read().strip().splitlines()[-1])
    print(t, d["ms_per_step"], d["verified_bit_exact_roundtrip"], {k:v for k,v in d["kernels_ms"].items() if v>0.2})
PY
