set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_c5_gputests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c5_bench.json 2> gpurun_out/r02_c5_bench.err
python tools/pip_microbench.py build/mb/*/pip3b.cubin --waves 8 > gpurun_out/r02_c5_mb3.log 2>&1
python tools/pip_microbench.py build/mb/*/pip2b.cubin --waves 8 > gpurun_out/r02_c5_mb2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_c5_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c5_ncu_bench.log 2>&1
