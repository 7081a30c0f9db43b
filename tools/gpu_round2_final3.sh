set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_f3_gputests.log
python bench.py > gpurun_out/r02_f3_bench.json 2> gpurun_out/r02_f3_bench.err
python bench.py --impl reference > gpurun_out/r02_f3_bench_ref.json 2> gpurun_out/r02_f3_bench_ref.err
python __graft_entry__.py smoke > gpurun_out/r02_f3_smoke.log 2>&1
cat gpurun_out/r02_f3_gputests.log gpurun_out/r02_f3_smoke.log; tail -c 600 gpurun_out/r02_f3_bench.err
