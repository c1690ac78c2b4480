#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
tail -c 1500 gpurun_out/bench_n$N.json
tail -3 gpurun_out/bench_n$N.err
