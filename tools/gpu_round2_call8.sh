set -x
for v in ship r4 n3; do
ncu --set full --clock-control none --import-source on -k regex:pip3b -s 3 -c 1 -o gpurun_out/r02_c8_pip3b_$v python tools/pip_microbench.py build/mb/$v/pip3b.cubin --waves 8 --reps 2 > gpurun_out/r02_c8_ncu_$v.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:pip2b -s 3 -c 1 -o gpurun_out/r02_c8_pip2b_ship python tools/pip_microbench.py build/mb/ship/pip2b.cubin --waves 8 --reps 2 > gpurun_out/r02_c8_ncu_2b.log 2>&1
python -m pytest tests/test_dispersion.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_c8_dispersion.log
