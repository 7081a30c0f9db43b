set -x
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_c33_$tag.json 2> gpurun_out/r02_c33_$tag.err; }
run default X=1
run tree NCCL_ALGO=Tree
run ring NCCL_ALGO=Ring
run ll128 NCCL_PROTO=LL128
