#!/usr/bin/env python
"""BASELINE config 5: ensemble inference (5 members + conformal heads, forward only) on 100 000 synthetic structures,
structures sharded across the ranks, every rank holding all members (reference ``predict.ensemble_predict``,
``scripts/predict.py:582-653``; member loop ``:604-623``; conformal interval ``train.py:1053-1076``).

    python scripts/bench_inference.py [--structures 100000] [--batch 256] [--dtypes bf16,fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
        scripts/bench_inference.py

Each rank keeps a device-resident store of ``--pool`` distinct 32-atom / 12-neighbour crystals (``DeviceGraphStore``) and
draws its share of the 100 000 structures from it batch by batch: graph ids go up (2 KB per batch), ``alignn_collate``
builds the PyG batch in HBM, ``EnsemblePredictor`` replays the five members + mixture moments as one CUDA graph, the
conformal interval + inverse log transform run as one more kernel (``alignn_ensemble_post``), and ``[B, 2]`` mean / lower /
upper come back to pinned host memory every batch.  Timed on the device (CUDA events), max over ranks; one JSON line per
dtype.  No collective on the data path: structures are independent."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import gnn_elasticity_predictor_b200 as pkg  # noqa: E402
from gnn_elasticity_predictor_b200 import dataset, dp, engine, ensemble  # noqa: E402
from gnn_elasticity_predictor_b200.synthetic import make_crystal  # noqa: E402
from bench import ARCH  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--structures", type=int, default=100_000)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--pool", type=int, default=1024)
    ap.add_argument("--members", type=int, default=5)
    ap.add_argument("--dtypes", default="bf16,fp32")
    ap.add_argument("--lg-inc", default="pyg", choices=["pyg", "bonds"])
    args = ap.parse_args()
    rank, local_rank, world = dp.init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    lo, hi = rank * args.structures // world, (rank + 1) * args.structures // world
    n_local = hi - lo
    n_batches = (n_local + args.batch - 1) // args.batch          # the last batch is drawn full; surplus rows are dropped
    gen = torch.Generator().manual_seed(777 + rank)
    store = dataset.DeviceGraphStore([make_crystal(32, 12, gen) for _ in range(args.pool)], dev, lg_inc=args.lg_inc)
    ids = [torch.randint(0, args.pool, (args.batch,), generator=gen).pin_memory() for _ in range(8)]
    ids_np = [t.numpy() for t in ids]
    ids_dev = [torch.empty(args.batch, dtype=torch.int64, device=dev) for _ in range(2)]
    members = []
    for m in range(args.members):
        torch.manual_seed(42 + 1007 * m)
        members.append(pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.15, **ARCH), 2).to(dev).eval())
    # conformal heads (artifacts/ensemble/conformal.pt: q scaled, alpha 0.1) + LogTransformer statistics (scaler_state.pt)
    q = torch.tensor([0.9173, 1.5967], device=dev)
    log_means, log_stds = torch.tensor([4.3228, 3.5567], device=dev), torch.tensor([0.9051, 0.9405], device=dev)
    out_host = [torch.empty(3, args.batch, 2).pin_memory() for _ in range(2)]

    for name in args.dtypes.split(","):
        cd = {"bf16": torch.bfloat16, "fp32": torch.float32}[name]
        pred = engine.EnsemblePredictor(members, compute_dtype=cd, graph=True, graph_warmup=1)
        probe = store.collate(ids_np[0])
        static = None

        def one(i):
            nonlocal static
            ids_dev[i % 2].copy_(ids[i % 8], non_blocking=True)
            b = store.collate(ids_np[i % 8], ids_device=ids_dev[i % 2], out=static)
            mean_z, var_z, std_z = pred.predict(b)
            cap = pred._captured.get(engine.TrainStep.signature(b))
            if static is None and cap is not None:
                static = cap[1]                      # collate straight into the captured graph's input buffers from now on
            # mixture moments are inside the graph; the conformal interval + inverse log transform are one more kernel
            # (a one-"member" call whose variance is the mixture variance reproduces the same moments)
            post = ensemble.ensemble_post(mean_z.unsqueeze(0), var_z.log().unsqueeze(0), min_logvar_floor=-80.0, q=q,
                                          method="scaled", log_means=log_means, log_stds=log_stds)
            out_host[i % 2].copy_(torch.stack([post["mean"], post["lower"], post["upper"]]), non_blocking=True)

        for i in range(6):
            one(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(n_batches):
            one(i)
        t1.record()
        torch.cuda.synchronize()
        t = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        if rank == 0:
            print(json.dumps({
                "metric": "ensemble inference structures/s (5 members + conformal heads, forward only)",
                "value": args.structures / (ms / 1e3), "unit": "structures/s", "n_gpus": world, "dtype": name,
                "structures": args.structures, "batch": args.batch, "batches_per_rank": n_batches,
                "ms_per_batch": ms / n_batches, "members": args.members, "lg_inc": args.lg_inc,
                "replays": pred.replays, "eager_calls": pred.eager_calls, "sizes": probe.sizes,
                "h2d_bytes_per_batch": args.batch * 8, "d2h_bytes_per_batch": out_host[0].numel() * 4,
                "how": "structures sharded over ranks, all members on every rank; ids H2D -> alignn_collate -> "
                       "EnsemblePredictor (one CUDA graph: plans + 5 member forwards + mixture moments) -> "
                       "alignn_ensemble_post (interval + inverse log transform) -> D2H; CUDA events, max over ranks"}),
                flush=True)
        del pred
    if world > 1:
        dp.shutdown()


if __name__ == "__main__":
    main()
