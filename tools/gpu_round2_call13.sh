set -x
python tools/pip_microbench.py build/mb/*/pip3b.cubin --waves 8 > gpurun_out/r02_c13_mb3.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_c13_gputests.log
python bench.py --steps 10 --warmup 3 --nve-steps 0 > gpurun_out/r02_c13_bench.json 2> gpurun_out/r02_c13_bench.err
cat gpurun_out/r02_c13_mb3.log gpurun_out/r02_c13_gputests.log
