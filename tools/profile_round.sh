#!/bin/bash
# One GPU call: the bench line without a profiler, then the ncu launch list and `--set full` captures of the kernels of
# a config-2 step (B200_PROFILING.md recipe), then the launch lists of a config-3 and a config-5 slice.
# Outputs under gpurun_out/; tools/ncu_summary.py turns them into profiles/.
set -u
tag=${1:-r02}
args="--no-e2e --no-cpu-baseline --configs= --no-lz"
python bench.py --steps 2 --warmup 3 $args > gpurun_out/${tag}_prof_bench.json 2> gpurun_out/${tag}_prof_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 $args > gpurun_out/${tag}_ncu_launch.log 2>&1
# (the merged-back directory is limited to 64 MiB: captures with source for the two dominant kernels, one without for the rest)
ncu --set full --import-source on --clock-control none -k regex:k_rans_decode_tiles_s0 -c 1 \
    -o gpurun_out/prof_${tag}_decode -f python bench.py --steps 1 --warmup 1 $args > gpurun_out/${tag}_ncu_full_a.log 2>&1
# the encoder's populated table-size class at config 2 is the second launch (rows <= 128)
ncu --set full --import-source on --clock-control none -k regex:k_rans_encode_ws --launch-skip 1 -c 1 \
    -o gpurun_out/prof_${tag}_encode -f python bench.py --steps 1 --warmup 1 $args > gpurun_out/${tag}_ncu_full_c.log 2>&1
ncu --set full --clock-control none \
    -k 'regex:k_rans_encode|k_tile_residuals_s0|k_gather_streams|k_build_tables|k_parse_streams' -c 8 \
    -o gpurun_out/prof_${tag}_rest -f python bench.py --steps 1 --warmup 1 $args > gpurun_out/${tag}_ncu_full_b.log 2>&1
python tools/bench_modes.py --images 64 --modes 2 --steps 2 --cpu-tiles 1 > gpurun_out/${tag}_modes_config3.json 2>&1
python tools/bench_modes.py --images 4096 --width 256 --height 256 --modes 4 --steps 2 --cpu-tiles 1 > gpurun_out/${tag}_modes_config5.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_launches_config3.csv \
    python tools/bench_modes.py --images 64 --modes 2 --steps 1 --cpu-tiles 1 > gpurun_out/${tag}_ncu_l3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_launches_config5.csv \
    python tools/bench_modes.py --images 4096 --width 256 --height 256 --modes 4 --steps 1 --cpu-tiles 1 > gpurun_out/${tag}_ncu_l5.log 2>&1
ls -la gpurun_out/prof_${tag}_*.ncu-rep gpurun_out/${tag}_launches*.csv
du -sh gpurun_out
