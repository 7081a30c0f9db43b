set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_c28_gputests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c28_bench.json 2> gpurun_out/r02_c28_bench.err
MBPOL_SERIAL_ORDERS=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --nve-steps 0 > gpurun_out/r02_c28_bench_serial.json 2> gpurun_out/r02_c28_bench_serial.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --nve-steps 0 --workload synthetic_32768 > gpurun_out/r02_c28_bench_32768.json 2>/dev/null
MBPOL_SERIAL_ORDERS=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --nve-steps 0 --workload synthetic_32768 > gpurun_out/r02_c28_bench_32768_serial.json 2>/dev/null
cat gpurun_out/r02_c28_gputests.log; tail -2 gpurun_out/r02_c28_bench.err
