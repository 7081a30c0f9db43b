set -x
nvidia-smi -L | wc -l
for wl in synthetic_262144 synthetic_32768; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --workload $wl --no-cpu-baseline > gpurun_out/r02_s8_${wl}_8gpu.json 2> gpurun_out/r02_s8_${wl}_8gpu.err
done
for wl in synthetic_262144 synthetic_32768; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 --workload $wl --no-cpu-baseline > gpurun_out/r02_s8_${wl}_4gpu.json 2> gpurun_out/r02_s8_${wl}_4gpu.err
done
tail -3 gpurun_out/r02_s8_*.err
cat gpurun_out/r02_s8_*.json
