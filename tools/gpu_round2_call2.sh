set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_c2_gputests.log
python tools/pip_microbench.py build/mb/*/pip3b.cubin --waves 8 > gpurun_out/r02_c2_mb3.log 2>&1
python tools/pip_microbench.py build/mb/*/pip2b.cubin --waves 8 > gpurun_out/r02_c2_mb2.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_c2_bench.json 2> gpurun_out/r02_c2_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_c2_bench_ref.json 2> gpurun_out/r02_c2_bench_ref.err
