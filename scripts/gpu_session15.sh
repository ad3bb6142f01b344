#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edgeattn.py tests/test_gpu_engine.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_v9.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/pytest_v9.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_plain.log 2>&1; echo "smoke $?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "
import __graft_entry__ as g
g.smoke()
import torch, gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import engine
torch.manual_seed(0)
m = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(206, 36, 11, 289, 2, 256, 2, 4, 0.1), 2).cuda(); m.base.compute_dtype = torch.bfloat16; m.train()
b = pkg.synthetic_batch(6, 16, 12, seed=1).to('cuda'); tz = pkg.zscore_targets(b.y, 6)
ts = engine.TrainStep(m, graph=False)
for _ in range(2): print(float(ts.step(b, tz)[0]))
" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"
tail -6 gpurun_out/sanitizer_memcheck.log | cut -c1-200
