set -x
python tools/pip_microbench.py build/mb/*/pip3b.cubin --waves 8 > gpurun_out/r02_c11_mb3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pip3b_eval -s 3 -c 1 -o gpurun_out/r02_c11_m6k16 python tools/pip_microbench.py build/mb/m6k16/pip3b.cubin --waves 8 > gpurun_out/r02_c11_ncu.log 2>&1
cat gpurun_out/r02_c11_mb3.log
