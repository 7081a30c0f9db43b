set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_c6_gputests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c6_bench.json 2> gpurun_out/r02_c6_bench.err
python tools/pip_microbench.py build/mb/*/pip3b.cubin --waves 8 > gpurun_out/r02_c6_mb3.log 2>&1
python tools/pip_microbench.py build/mb/*/pip2b.cubin --waves 8 > gpurun_out/r02_c6_mb2.log 2>&1
