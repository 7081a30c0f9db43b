set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_f2_gputests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_f2_bench.json 2> gpurun_out/r02_f2_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_f2_bench_ref.json 2> gpurun_out/r02_f2_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r02_f2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --nve-steps 0 > gpurun_out/r02_f2_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pip3b_eval -s 3 -c 1 -o gpurun_out/r02_f2_pip3b python bench.py --steps 1 --warmup 3 --no-cpu-baseline --nve-steps 0 > gpurun_out/r02_f2_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pip2b_eval -s 3 -c 1 -o gpurun_out/r02_f2_pip2b python bench.py --steps 1 --warmup 3 --no-cpu-baseline --nve-steps 0 > gpurun_out/r02_f2_ncu2.log 2>&1
python __graft_entry__.py smoke > gpurun_out/r02_f2_smoke.log 2>&1
cat gpurun_out/r02_f2_gputests.log gpurun_out/r02_f2_smoke.log
