#!/bin/bash
# Everything that needs more than one GPU, in ONE gpurun call:  gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_multi.sh 8'
# (N = GPUs on the box; the N=2 / N=4 points run on a subset).  Outputs under gpurun_out/r02_multi/.
N=${1:-8}
OUT=gpurun_out/r02_multi
mkdir -p $OUT
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "${@:2}"; }
# 1. N-GPU == 1-GPU on the real engine (NCCL all-reduce inside the step graph)
for dt in fp32 bf16; do
  run 2 scripts/check_dp_equality.py --dtype $dt --steps 4 > $OUT/dp_equality_$dt.txt 2>&1; echo "rc=$?" >> $OUT/dp_equality_$dt.txt
done
# 2. weak scaling (256 graphs / GPU) and strong scaling (config 3 as stated: global 2 048)
for n in 2 4 8; do
  [ $n -le $N ] || continue
  run $n bench.py --gpus $n --steps 20 --warmup 5 --no-bonds > $OUT/bench_weak_n$n.json 2> $OUT/bench_weak_n$n.err
  run $n bench.py --gpus $n --steps 20 --warmup 5 --scaling strong --no-e2e > $OUT/bench_strong_n$n.json 2> $OUT/bench_strong_n$n.err
done
timeout 420 python bench.py --steps 10 --warmup 3 --scaling strong --no-e2e --no-cpu-baseline > $OUT/bench_strong_n1.json 2> $OUT/bench_strong_n1.err
# 3. member-per-GPU placement: 5 members on 5 GPUs, no exchange
if [ $N -ge 5 ]; then
  run 5 bench.py --gpus 5 --steps 20 --warmup 5 --placement members --no-bonds > $OUT/bench_members_n5.json 2> $OUT/bench_members_n5.err
fi
# 4. config 5: 100 000 structures, 5 members + conformal heads, sharded over the GPUs (bf16 and fp32)
run $N scripts/bench_inference.py > $OUT/inference_n$N.json 2> $OUT/inference_n$N.err
timeout 300 python scripts/bench_inference.py --structures 20000 > $OUT/inference_n1.json 2> $OUT/inference_n1.err
tail -c 600 $OUT/*.json | cut -c1-400
