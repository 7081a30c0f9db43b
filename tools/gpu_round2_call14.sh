set -x
python tools/pip_microbench.py build/mb/*/pip3b.cubin --waves 8 > gpurun_out/r02_c14_mb3.log 2>&1
cat gpurun_out/r02_c14_mb3.log
