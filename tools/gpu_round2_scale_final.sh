set -x
N=$(nvidia-smi -L | wc -l)
for n in 8 4 2; do
for wl in synthetic_262144 synthetic_32768; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 --workload $wl --no-cpu-baseline > gpurun_out/r02_sf_${wl}_${n}gpu.json 2> gpurun_out/r02_sf_${wl}_${n}gpu.err
done
done
tail -2 gpurun_out/r02_sf_*.err
