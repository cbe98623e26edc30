#!/bin/bash
set -u
args="--images 256 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-lz --configs=3,5"
for i in 1 2; do
timeout 300 python bench.py $args > gpurun_out/pa_fused$i.json 2> gpurun_out/pa_fused$i.err
done
python - <<'PY'
import json
for t in ("fused1","fused2"):
    d=json.loads(open(f"gpurun_out/pa_{t}.json").read().strip().splitlines()[-1])
    for k,v in d["configs"].items():
        print(t,k,v["encode_ms_calls"],v["decode_ms_calls"])
PY
