#!/bin/bash
# N-GPU check: sharded parity tests, then the bench at N
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n$N.txt
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/test_sharded_n$N.log 2>&1
echo "sharded tests rc=$? $(tail -1 gpurun_out/test_sharded_n$N.log)"
grep -E "FAILED|Error|error|assert" gpurun_out/test_sharded_n$N.log | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
    print(json.dumps(d.get("sharded"), indent=1)[:3000])
    print({k: d[k] for k in ("value", "ms_per_step", "e2e")})
except Exception as e:
    print("parse failed", e)
PY
tail -5 gpurun_out/bench_n$N.err
