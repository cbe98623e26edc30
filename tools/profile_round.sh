#!/bin/bash
# One GPU call: the bench line without a profiler, then the ncu launch list and one `--set full` capture of the kernels of
# a config-2 step (B200_PROFILING.md recipe).  Outputs under gpurun_out/; tools/ncu_summary.py turns them into profiles/.
set -u
tag=${1:-r02}
args="--no-e2e --no-cpu-baseline --configs= --no-lz"
python bench.py --steps 2 --warmup 3 $args > gpurun_out/${tag}_prof_bench.json 2> gpurun_out/${tag}_prof_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 $args > gpurun_out/${tag}_ncu_launch.log 2>&1
# (the merged-back directory is limited to 64 MiB: one capture with source for the dominant kernel, one without for the rest)
ncu --set full --import-source on --clock-control none -k regex:k_rans_decode_tiles_s0 -c 1 \
    -o gpurun_out/prof_${tag}_decode -f python bench.py --steps 1 --warmup 1 $args > gpurun_out/${tag}_ncu_full_a.log 2>&1
ncu --set full --clock-control none \
    -k 'regex:k_rans_encode|k_tile_residuals_s0|k_gather_streams|k_build_tables|k_parse_streams' -c 8 \
    -o gpurun_out/prof_${tag}_rest -f python bench.py --steps 1 --warmup 1 $args > gpurun_out/${tag}_ncu_full_b.log 2>&1
ls -la gpurun_out/prof_${tag}_*.ncu-rep gpurun_out/${tag}_launches.csv
du -sh gpurun_out
