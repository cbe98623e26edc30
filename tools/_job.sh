python -m pytest tests -m gpu -x -q > gpurun_out/t11.log 2>&1; tail -5 gpurun_out/t11.log
python bench.py --no-cpu-baseline > gpurun_out/b11.json 2> gpurun_out/b11.err; tail -c 300 gpurun_out/b11.err
