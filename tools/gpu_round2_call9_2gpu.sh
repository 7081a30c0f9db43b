set -x
nvidia-smi -L
gcc -std=c99 -Wall examples/capi_demo_2gpu.c -Iinclude -I/usr/local/cuda/include -Lmbpol_openmm_plugin_b200 -lmbpol_b200 -L/usr/local/cuda/lib64 -lcudart -lpthread -lm -Wl,-rpath,$PWD/mbpol_openmm_plugin_b200 -o /tmp/capi_demo_2gpu && /tmp/capi_demo_2gpu > gpurun_out/r02_c9_capi_demo_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c9_capi_demo_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --workload synthetic_32768 > gpurun_out/r02_c9_bench_32768_2gpu.json 2> gpurun_out/r02_c9_bench_32768_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_c9_bench_262144_2gpu.json 2> gpurun_out/r02_c9_bench_262144_2gpu.err
tail -5 gpurun_out/*2gpu.err
