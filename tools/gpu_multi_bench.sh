#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n${N}_late.json 2> gpurun_out/bench_n${N}_late.err; echo "bench n$N rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n${N}_late.json").read().strip().splitlines()[-1])
    print(json.dumps(d.get("sharded"), indent=1)[:2500])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")})
except Exception as e:
    print("parse failed", e)
PY
tail -3 gpurun_out/bench_n${N}_late.err
