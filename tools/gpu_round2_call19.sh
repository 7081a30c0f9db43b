set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_c19_gputests.log
python bench.py --steps 10 --warmup 3 --nve-steps 0 > gpurun_out/r02_c19_bench.json 2> gpurun_out/r02_c19_bench.err
cat gpurun_out/r02_c19_gputests.log
