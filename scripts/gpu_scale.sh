#!/bin/bash
# usage: gpu_scale.sh N   -- the driver's multi-GPU launch line for bench.py
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_v12_n$N.json 2> gpurun_out/bench_v12_n$N.err; echo "bench N=$N exit $?"
tail -3 gpurun_out/bench_v12_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_v12_n$N.json").read().strip().splitlines()[-1])
    print("N=$N", {k: d.get(k) for k in ("value", "ms_per_step", "n_gpus")}, "e2e", (d.get("e2e") or {}).get("value"), "store", (d.get("e2e_device_store") or {}).get("value"))
except Exception as e:
    print("ERR", e)
PY
