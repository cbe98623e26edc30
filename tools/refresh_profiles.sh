#!/bin/bash
# Turn what tools/final_job.sh / tools/profile_round.sh left under gpurun_out/ (tag r02d) into the tables under profiles/.
set -eu
tag=${1:-r02d}
for r in decode encode rest; do ncu -i gpurun_out/prof_${tag}_$r.ncu-rep --page raw --csv > /tmp/raw_$r.csv 2>/dev/null; done
python - <<'PY'
import csv
rows_all = None
for r in ("decode", "encode", "rest"):
    rows = list(csv.reader(open(f"/tmp/raw_{r}.csv")))
    if rows_all is None:
        rows_all, hdr = rows, rows[0]
    else:
        idx = {n: i for i, n in enumerate(rows[0])}
        for row in rows[2:]:
            rows_all.append([row[idx[n]] if n in idx and idx[n] < len(row) else "" for n in hdr])
csv.writer(open("/tmp/raw_all.csv", "w")).writerows(rows_all)
PY
python tools/ncu_summary.py full /tmp/raw_all.csv --traffic profiles/ncu_traffic.json > profiles/r02_ncu_full_table.md
cp gpurun_out/${tag}_launches.csv profiles/r02_launches.csv
cp gpurun_out/${tag}_launches_config3.csv profiles/r02_launches_config3_64frames_mode2.csv
cp gpurun_out/${tag}_launches_config5.csv profiles/r02_launches_config5_4096thumbs_mode4.csv
python tools/ncu_summary.py launches profiles/r02_launches.csv > profiles/r02_launch_table.md
cp gpurun_out/${tag}_prof_bench.json profiles/r02_bench_profiled_build.json
tail -1 gpurun_out/${tag}_bench.json > profiles/r02_bench_1gpu.json
tail -1 gpurun_out/${tag}_bench_ref.json > profiles/r02_bench_1gpu_reference_arm.json
cp gpurun_out/${tag}_gpu_tests.log profiles/r02_gpu_tests.log
