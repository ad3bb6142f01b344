run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "${@:2}"; }
N=${1:-2}
mkdir -p gpurun_out/r02_multi2
for rep in 1; do
for mode in early late; do
  flag=""; [ $mode = early ] && flag="--early-allreduce"
  run $N bench.py --gpus $N --steps 40 --warmup 5 --no-bonds --no-e2e $flag > gpurun_out/r02_multi2/bench_n${N}_$mode.json 2> gpurun_out/r02_multi2/bench_n${N}_$mode.err
  python -c "
import json;d=json.loads(open('gpurun_out/r02_multi2/bench_n${N}_$mode.json').read().strip().splitlines()[-1]);print('$mode', d['value'], d['ms_per_step'])"
done; done
timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-bonds --no-e2e > gpurun_out/r02_multi2/bench_n1.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r02_multi2/bench_n1.json').read().strip().splitlines()[-1]);print('n1', d['value'], d['ms_per_step'])"
