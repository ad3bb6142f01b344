#!/bin/bash
# final-code refresh of the 8-GPU lines: weak scaling N=8 and config-5 inference N=8 (bf16 + fp32)
N=${1:-8}
OUT=gpurun_out/r02_multi_final
mkdir -p $OUT
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "${@:2}"; }
run $N bench.py --gpus $N --steps 20 --warmup 5 --no-bonds > $OUT/bench_weak_n$N.json 2> $OUT/bench_weak_n$N.err
run $N scripts/bench_inference.py > $OUT/inference_n$N.json 2> $OUT/inference_n$N.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-bonds --no-cpu-baseline > $OUT/bench_n1_same_box.json 2> $OUT/bench_n1.err
tail -c 400 $OUT/*.json | cut -c1-300
