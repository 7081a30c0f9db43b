set -x
python tools/pip_microbench.py build/mb/*/pip3b.cubin --waves 8 > gpurun_out/r02_c4_mb3.log 2>&1
python -m pytest tests/test_openmm_adapter.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_c4_adapter.log
ncu --set full --clock-control none --import-source on -k regex:forces_kernel -c 2 -o gpurun_out/r02_c4_forces python bench.py --workload synthetic_32768 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c4_ncu.log 2>&1
