set -x
python tools/pip_microbench.py build/mb/*/pip3b.cubin --waves 8 > gpurun_out/r02_c10_mb3.log 2>&1
python tools/pip_microbench.py build/mb/*/pip2b.cubin --waves 8 > gpurun_out/r02_c10_mb2.log 2>&1
